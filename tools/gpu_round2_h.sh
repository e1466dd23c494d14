#!/bin/bash
# Round 2: the whole GPU suite on the final kernels (C4 oracle on the strict-fp32 GPU executor), the default bench line.
TAG=${1:-r2h}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rf --durations=10 > gpurun_out/pytest_${TAG}.log 2>&1
tail -25 gpurun_out/pytest_${TAG}.log
python bench.py > gpurun_out/bench_default_${TAG}.json 2> gpurun_out/bench_default_${TAG}.err
cat gpurun_out/bench_default_${TAG}.json; tail -3 gpurun_out/bench_default_${TAG}.err
python bench.py --impl cudnn --steps 5 --warmup 2 > gpurun_out/bench_cudnn_${TAG}.json 2> gpurun_out/bench_cudnn_${TAG}.err
cat gpurun_out/bench_cudnn_${TAG}.json; tail -3 gpurun_out/bench_cudnn_${TAG}.err
python __graft_entry__.py smoke 2>&1 | tail -2
