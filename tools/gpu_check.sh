#!/bin/bash
# One round-trip on the GPU box: parity tests, bench lines, ncu launch list.  Usage (from the repo root):
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_${TAG}.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cat gpurun_out/bench_b4_${TAG}.json
python bench.py --steps 5 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/bench_b1_${TAG}.json 2> gpurun_out/bench_b1_${TAG}.err
cat gpurun_out/bench_b1_${TAG}.json
python tools/conv_layer_bench.py 256 1 > gpurun_out/convbench_b1_${TAG}.log 2>&1
cat gpurun_out/convbench_b1_${TAG}.log
python tools/aux_bench.py 256 1 > gpurun_out/auxbench_b1_${TAG}.log 2>&1
cat gpurun_out/auxbench_b1_${TAG}.log
