#!/bin/bash
# Final evidence of round 2 on the final tree: whole GPU suite, the default bench line (all legs), then the long run with
# its clock trace, the other BASELINE configurations, and the single-pass ncu captures (only ncu in this call).
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_gpu_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_gpu_${TAG}.log
python bench.py > gpurun_out/bench_default_${TAG}.json 2> gpurun_out/bench_default_${TAG}.err
cut -c1-200 gpurun_out/bench_default_${TAG}.json; tail -2 gpurun_out/bench_default_${TAG}.err
python bench.py --impl cudnn > gpurun_out/bench_cudnn_${TAG}.json 2> gpurun_out/bench_cudnn_${TAG}.err
cut -c1-300 gpurun_out/bench_cudnn_${TAG}.json
bash tools/gpu_round2_i.sh ${TAG}
