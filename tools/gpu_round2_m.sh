#!/bin/bash
TAG=${1:-r2m}
mkdir -p gpurun_out
python -m pytest tests/test_aux_gpu.py tests/test_guard_gpu.py tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py -m gpu -q -rf -k "not properties" > gpurun_out/pytest_${TAG}.log 2>&1
tail -6 gpurun_out/pytest_${TAG}.log
python tools/aux_bench.py 256 1 2>&1 | grep "K7\|K2 stem\|K8" | tee gpurun_out/auxbench_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cut -c1-330 gpurun_out/bench_b4_${TAG}.json; tail -3 gpurun_out/bench_b4_${TAG}.err
