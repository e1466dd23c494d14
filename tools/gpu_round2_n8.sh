#!/bin/bash
TAG=${1:-r2n8}
N=${2:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err
cat gpurun_out/bench_n${N}_${TAG}.json | cut -c1-1500; tail -3 gpurun_out/bench_n${N}_${TAG}.err
