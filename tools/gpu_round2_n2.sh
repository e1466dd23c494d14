#!/bin/bash
# Two GPUs: the bench under torchrun exactly as the driver launches it (ours, reference arm), and the training line.
TAG=${1:-r2n2}
N=${2:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err
cat gpurun_out/bench_n${N}_${TAG}.json | cut -c1-1200; tail -3 gpurun_out/bench_n${N}_${TAG}.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus $N --steps 2 --warmup 1 \
    > gpurun_out/bench_ref_n${N}_${TAG}.json 2> gpurun_out/bench_ref_n${N}_${TAG}.err
cat gpurun_out/bench_ref_n${N}_${TAG}.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --mode train --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/bench_train_n${N}_${TAG}.json 2> gpurun_out/bench_train_n${N}_${TAG}.err
cat gpurun_out/bench_train_n${N}_${TAG}.json | cut -c1-900; tail -3 gpurun_out/bench_train_n${N}_${TAG}.err
