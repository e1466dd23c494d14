#!/bin/bash
# Round 2: plane-ring kernel for Cout 128 + residual prefetch: parity, timings.
TAG=${1:-r2j}
mkdir -p gpurun_out
python -m pytest tests/test_conv3d_gpu.py tests/test_model_gpu.py tests/test_guard_gpu.py tests/test_training_gpu.py tests/test_backward_gpu.py tests/test_fullsize_gpu.py \
    -m gpu -q -rf --durations=5 -k "not properties" > gpurun_out/pytest_${TAG}.log 2>&1
tail -12 gpurun_out/pytest_${TAG}.log
python tools/conv_layer_bench.py 256 1 > gpurun_out/convbench_b1_${TAG}.log 2>&1
cat gpurun_out/convbench_b1_${TAG}.log
DRAM_B200_SLAB128=0 python tools/conv_layer_bench.py 256 1 auto "layer2 128" 2>&1 | grep layer2
python tools/conv_one.py 4 32 32 32 128 0 128 3 1 1 128
DRAM_B200_SLAB128=0 python tools/conv_one.py 4 32 32 32 128 0 128 3 1 1 128
python tools/conv_one.py 4 64 64 64 64 0 64 3 1 1 64
python tools/engine_profile.py med3ddram 256,256,256 4 > gpurun_out/engine_b4_${TAG}.log 2>&1
grep -v "layer3\.[1-5]\|layer4\.[12]\|layer2\.[23]\|layer1\.[12]" gpurun_out/engine_b4_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cut -c1-330 gpurun_out/bench_b4_${TAG}.json; tail -3 gpurun_out/bench_b4_${TAG}.err
python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline --no-yardstick --no-train-field > gpurun_out/bench_b1_${TAG}.json 2> gpurun_out/bench_b1_${TAG}.err
cut -c1-330 gpurun_out/bench_b1_${TAG}.json
python bench.py --steps 5 --warmup 3 --batch 1 --arch med3ddram50 --dims 400,512,512 --no-cpu-baseline --no-yardstick --no-train-field > gpurun_out/bench_c4_${TAG}.json 2> gpurun_out/bench_c4_${TAG}.err
cut -c1-330 gpurun_out/bench_c4_${TAG}.json
