#!/bin/bash
# Round 2, first box: the whole GPU suite (no -x: every failure listed), bench at batch 4 and batch 1, per-layer timings.
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; free -g | head -2; nproc
python -m pytest tests -m gpu -q -rf -s --durations=15 > gpurun_out/pytest_${TAG}.log 2>&1
tail -60 gpurun_out/pytest_${TAG}.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cat gpurun_out/bench_b4_${TAG}.json; tail -3 gpurun_out/bench_b4_${TAG}.err
DRAM_B200_GRAPH=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b4_nograph_${TAG}.json 2> gpurun_out/bench_b4_nograph_${TAG}.err
cat gpurun_out/bench_b4_nograph_${TAG}.json
python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/bench_b1_${TAG}.json 2> gpurun_out/bench_b1_${TAG}.err
cat gpurun_out/bench_b1_${TAG}.json
DRAM_B200_GRAPH=0 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/bench_b1_nograph_${TAG}.json 2> gpurun_out/bench_b1_nograph_${TAG}.err
cat gpurun_out/bench_b1_nograph_${TAG}.json
python tools/conv_layer_bench.py 256 1 > gpurun_out/convbench_b1_${TAG}.log 2>&1
cat gpurun_out/convbench_b1_${TAG}.log
python tools/aux_bench.py 256 1 > gpurun_out/auxbench_b1_${TAG}.log 2>&1
cat gpurun_out/auxbench_b1_${TAG}.log
