#!/bin/bash
# Round 2, fourth box: stem / epilogue / K7 / K13-GEMM rework: parity, kernel timings, bench, and ncu --set full source
# captures of the three kernels still furthest from their roofline (us3 = slab<32>, layer2 = umma<128>, K13 gather).
TAG=${1:-r2d}
mkdir -p gpurun_out
python -m pytest tests/test_conv3d_gpu.py tests/test_model_gpu.py tests/test_aux_gpu.py tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py \
    -m gpu -q -rf --durations=5 -k "not c2_ and not properties and not c4_" > gpurun_out/pytest_${TAG}.log 2>&1
tail -12 gpurun_out/pytest_${TAG}.log
python tools/aux_bench.py 256 1 > gpurun_out/auxbench_b1_${TAG}.log 2>&1
cat gpurun_out/auxbench_b1_${TAG}.log
python tools/engine_profile.py med3ddram 256,256,256 4 > gpurun_out/engine_b4_${TAG}.log 2>&1
cat gpurun_out/engine_b4_${TAG}.log
python tools/engine_profile.py med3ddram 256,256,256 1 > gpurun_out/engine_b1_${TAG}.log 2>&1
cat gpurun_out/engine_b1_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cat gpurun_out/bench_b4_${TAG}.json; tail -3 gpurun_out/bench_b4_${TAG}.err
python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline --no-yardstick > gpurun_out/bench_b1_${TAG}.json 2> gpurun_out/bench_b1_${TAG}.err
cat gpurun_out/bench_b1_${TAG}.json; tail -3 gpurun_out/bench_b1_${TAG}.err
# source-level captures (one launch each)
US3="python tools/conv_one.py 1 128 128 128 64 0 32 3 1 1"
L2="python tools/conv_one.py 4 32 32 32 128 0 128 3 1 1"
$US3 > gpurun_out/us3_plain_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv3d_slab' -s 4 -c 1 -o /tmp/us3_${TAG} $US3 > gpurun_out/ncu_us3_${TAG}.log 2>&1
ncu -i /tmp/us3_${TAG}.ncu-rep --page details > gpurun_out/us3_${TAG}.details.txt 2>&1
ncu -i /tmp/us3_${TAG}.ncu-rep --page source --csv > gpurun_out/us3_${TAG}.src.csv 2>&1
$L2 > gpurun_out/l2_plain_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv3d_umma' -s 4 -c 1 -o /tmp/l2_${TAG} $L2 > gpurun_out/ncu_l2_${TAG}.log 2>&1
ncu -i /tmp/l2_${TAG}.ncu-rep --page details > gpurun_out/l2_${TAG}.details.txt 2>&1
STEM="python tools/aux_bench.py 256 1"
ncu --set full --clock-control none --import-source on -k regex:'conv3d_stem_kernel|dram_upsample_mask_lean' -c 2 -o /tmp/stem_${TAG} $STEM > gpurun_out/ncu_stem_${TAG}.log 2>&1
ncu -i /tmp/stem_${TAG}.ncu-rep --page details > gpurun_out/stem_${TAG}.details.txt 2>&1
ls -la gpurun_out | tail -8
