#!/bin/bash
# Short single-GPU round-trip: the GPU suite, smoke, the training step timing and both bench lines.
#   gpurun --timeout 700 -- 'bash tools/gpu_check_quick.sh TAG'
TAG=${1:-r1n}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_${TAG}.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke_${TAG}.log
timeout 120 python tools/train_step_bench.py 256 1 resnet34segreg native native > gpurun_out/trainbench_native_${TAG}.log 2>&1
head -4 gpurun_out/trainbench_native_${TAG}.log
timeout 150 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_n1_${TAG}.json 2> gpurun_out/bench_train_n1_${TAG}.err
cut -c1-300 gpurun_out/bench_train_n1_${TAG}.json
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cut -c1-300 gpurun_out/bench_b4_${TAG}.json
