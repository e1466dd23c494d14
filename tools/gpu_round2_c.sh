#!/bin/bash
# Round 2, third box: plane-ring issuer rewrite + K13 (commuted us1.0): parity, per-layer tables at batch 1 and 4, A/B of
# us1 routes, bench lines, C4 (ResNet-50 400x512x512) bench.
TAG=${1:-r2c}
mkdir -p gpurun_out
python -m pytest tests/test_conv3d_gpu.py tests/test_model_gpu.py tests/test_aux_gpu.py tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py \
    -m gpu -q -rf -s --durations=8 -k "not c2_ and not properties" > gpurun_out/pytest_${TAG}.log 2>&1
tail -25 gpurun_out/pytest_${TAG}.log
python tools/conv_layer_bench.py 256 1 > gpurun_out/convbench_b1_${TAG}.log 2>&1
cat gpurun_out/convbench_b1_${TAG}.log
python tools/engine_profile.py med3ddram 256,256,256 4 > gpurun_out/engine_b4_${TAG}.log 2>&1
cat gpurun_out/engine_b4_${TAG}.log
DRAM_B200_US1=direct python tools/engine_profile.py med3ddram 256,256,256 4 2>&1 | grep -i "us1\|engine step" > gpurun_out/engine_b4_direct_${TAG}.log
cat gpurun_out/engine_b4_direct_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cat gpurun_out/bench_b4_${TAG}.json; tail -3 gpurun_out/bench_b4_${TAG}.err
python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline --no-yardstick > gpurun_out/bench_b1_${TAG}.json 2> gpurun_out/bench_b1_${TAG}.err
cat gpurun_out/bench_b1_${TAG}.json; tail -3 gpurun_out/bench_b1_${TAG}.err
python tools/engine_profile.py med3ddram50 400,512,512 1 > gpurun_out/engine_c4_${TAG}.log 2>&1
cat gpurun_out/engine_c4_${TAG}.log
python bench.py --steps 5 --warmup 3 --batch 1 --arch med3ddram50 --dims 400,512,512 --no-cpu-baseline --no-yardstick > gpurun_out/bench_c4_${TAG}.json 2> gpurun_out/bench_c4_${TAG}.err
cat gpurun_out/bench_c4_${TAG}.json; tail -3 gpurun_out/bench_c4_${TAG}.err
