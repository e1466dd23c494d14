#!/bin/bash
# Round 2, second box: the tests that failed or changed, bench with the new fields, per-layer table at batch 4, and the
# single-pass ncu captures (tensor pipe per conv launch, launch list, memory-bound kernels).
TAG=${1:-r2b}
mkdir -p gpurun_out
python -m pytest tests/test_fullsize_gpu.py tests/test_model_gpu.py tests/test_aux_gpu.py tests/test_pipeline_gpu.py \
    -m gpu -q -rf -s --durations=8 -k "c3_ or c4_resnet50_400x512x512_matches or refinit or stem_from_hu or window or from_hu or staged or saturation" \
    > gpurun_out/pytest_${TAG}.log 2>&1
tail -25 gpurun_out/pytest_${TAG}.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cat gpurun_out/bench_b4_${TAG}.json; tail -3 gpurun_out/bench_b4_${TAG}.err
python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline --no-yardstick > gpurun_out/bench_b1_${TAG}.json 2> gpurun_out/bench_b1_${TAG}.err
cat gpurun_out/bench_b1_${TAG}.json; tail -3 gpurun_out/bench_b1_${TAG}.err
python tools/engine_profile.py med3ddram 256,256,256 4 > gpurun_out/engine_b4_${TAG}.log 2>&1
cat gpurun_out/engine_b4_${TAG}.log
bash tools/gpu_profile_pipe.sh ${TAG}
