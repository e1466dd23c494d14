#!/bin/bash
# 2-GPU round-trip for the peer-memory SyncBatchNorm exchange (K10x):  gpurun --gpus 2 --timeout 700 -- 'bash tools/gpu_check_peer.sh TAG'
TAG=${1:-r1m}
N=${2:-2}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_backward_gpu.py -m gpu -x -q -k "peer_allreduce" 2>&1 | tail -8 | tee gpurun_out/pytest_peer_${TAG}.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/train_ddp_smoke.py 64 > gpurun_out/train_ddp_n${N}_${TAG}.log 2>&1
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/train_ddp_n${N}_${TAG}.log | tail -16
for MODE in nccl peer; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --mode train --steps 10 --warmup 3 --sync-bn $MODE > gpurun_out/bench_train_n${N}_${MODE}_${TAG}.json 2> gpurun_out/bench_train_n${N}_${MODE}_${TAG}.err
  cut -c1-330 gpurun_out/bench_train_n${N}_${MODE}_${TAG}.json; tail -2 gpurun_out/bench_train_n${N}_${MODE}_${TAG}.err
done
