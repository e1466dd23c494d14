#!/bin/bash
# Round 2: stem TMA-store epilogue, K13 gather/GEMM rework, K7 zero fill: parity + timings.
TAG=${1:-r2f}
mkdir -p gpurun_out
python -m pytest tests/test_conv3d_gpu.py tests/test_model_gpu.py tests/test_aux_gpu.py tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py tests/test_training_gpu.py \
    -m gpu -q -rf --durations=5 -k "not c2_ and not properties" > gpurun_out/pytest_${TAG}.log 2>&1
tail -12 gpurun_out/pytest_${TAG}.log
python tools/aux_bench.py 256 1 > gpurun_out/auxbench_b1_${TAG}.log 2>&1
grep "K2 stem\|K7\|K8" gpurun_out/auxbench_b1_${TAG}.log
python tools/engine_profile.py med3ddram 256,256,256 4 > gpurun_out/engine_b4_${TAG}.log 2>&1
grep -v "layer3\.[1-5]\|layer4\.[12]\|layer2\.[123]\|layer1\.[12]" gpurun_out/engine_b4_${TAG}.log
DRAM_B200_US1_EPILOGUE=direct python tools/engine_profile.py med3ddram 256,256,256 4 2>&1 | grep "us1.0.z\|engine step"
python tools/engine_profile.py med3ddram50 400,512,512 1 > gpurun_out/engine_c4_${TAG}.log 2>&1
grep "us1\|conv1 \|engine step" gpurun_out/engine_c4_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cut -c1-330 gpurun_out/bench_b4_${TAG}.json; tail -3 gpurun_out/bench_b4_${TAG}.err
python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline --no-yardstick > gpurun_out/bench_b1_${TAG}.json 2> gpurun_out/bench_b1_${TAG}.err
cut -c1-330 gpurun_out/bench_b1_${TAG}.json; tail -3 gpurun_out/bench_b1_${TAG}.err
python bench.py --steps 5 --warmup 3 --batch 1 --arch med3ddram50 --dims 400,512,512 --no-cpu-baseline --no-yardstick > gpurun_out/bench_c4_${TAG}.json 2> gpurun_out/bench_c4_${TAG}.err
cut -c1-330 gpurun_out/bench_c4_${TAG}.json; tail -3 gpurun_out/bench_c4_${TAG}.err
