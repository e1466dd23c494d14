#!/bin/bash
# Round 2: interleaved MMA order back on: reproducibility + parity + bench.
TAG=${1:-r2l}
mkdir -p gpurun_out
python tools/train_repro_check.py 3 2>&1 | tail -5
python -m pytest tests/test_conv3d_gpu.py tests/test_model_gpu.py tests/test_training_gpu.py tests/test_fullsize_gpu.py \
    -m gpu -q -rf --durations=5 -k "not properties" > gpurun_out/pytest_${TAG}.log 2>&1
tail -8 gpurun_out/pytest_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cut -c1-330 gpurun_out/bench_b4_${TAG}.json; tail -3 gpurun_out/bench_b4_${TAG}.err
python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline --no-yardstick --no-train-field > gpurun_out/bench_b1_${TAG}.json 2> gpurun_out/bench_b1_${TAG}.err
cut -c1-330 gpurun_out/bench_b1_${TAG}.json
