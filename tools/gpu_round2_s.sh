#!/bin/bash
# Final evidence of round 2, part 1: whole GPU suite + the default bench line (all legs) on the final tree.
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_gpu_${TAG}.log 2>&1
tail -5 gpurun_out/pytest_gpu_${TAG}.log
python bench.py > gpurun_out/bench_default_${TAG}.json 2> gpurun_out/bench_default_${TAG}.err
cut -c1-400 gpurun_out/bench_default_${TAG}.json; tail -2 gpurun_out/bench_default_${TAG}.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
