#!/bin/bash
# ncu evidence for profiles/: launch list of one bench run + full captures of the conv kernels and the
# memory-bound kernels.  Run only after the same bench command has exited 0 without ncu.
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh TAG'
TAG=${1:-r1}
BENCH="python bench.py --steps 2 --warmup 3 --batch 1 --no-cpu-baseline"
mkdir -p gpurun_out
$BENCH > gpurun_out/prof_bench_${TAG}.json 2> gpurun_out/prof_bench_${TAG}.err || exit 1
cat gpurun_out/prof_bench_${TAG}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
# one full step of conv launches (the last timed step): skip the warm-up + first timed step
ncu --set full --clock-control none --import-source on -k regex:'conv3d_' -s 156 -c 39 -o gpurun_out/conv_${TAG} \
    $BENCH > gpurun_out/ncu_conv_${TAG}.log 2>&1
ncu -i gpurun_out/conv_${TAG}.ncu-rep --page raw --csv > gpurun_out/conv_${TAG}.raw.csv
ncu --set full --clock-control none --import-source on \
    -k regex:'upsample2x|maxpool3d|dram_upsample_mask|window_|masked_pool_partial' -s 36 -c 9 -o gpurun_out/aux_${TAG} \
    $BENCH > gpurun_out/ncu_aux_${TAG}.log 2>&1
ncu -i gpurun_out/aux_${TAG}.ncu-rep --page raw --csv > gpurun_out/aux_${TAG}.raw.csv
ls -la gpurun_out | tail -12
