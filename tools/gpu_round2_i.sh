#!/bin/bash
# Round 2 evidence: a >= 5 s timed region with its clock trace, then the single-pass ncu captures (only ncu in this call).
TAG=${1:-r2i}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 200 > gpurun_out/clocks_long_${TAG}.csv &
SMI=$!
python bench.py --steps 250 --warmup 5 --no-cpu-baseline --no-yardstick --no-train-field > gpurun_out/bench_b4_long_${TAG}.json 2> gpurun_out/bench_b4_long_${TAG}.err
kill $SMI
cut -c1-600 gpurun_out/bench_b4_long_${TAG}.json; tail -2 gpurun_out/bench_b4_long_${TAG}.err
python bench.py --steps 30 --warmup 3 --batch 1 --no-cpu-baseline --no-yardstick --no-train-field > gpurun_out/bench_b1_${TAG}.json 2> gpurun_out/bench_b1_${TAG}.err
cut -c1-400 gpurun_out/bench_b1_${TAG}.json
python bench.py --steps 10 --warmup 3 --arch med3ddram18 --no-cpu-baseline --no-yardstick --no-train-field > gpurun_out/bench_c2_${TAG}.json 2> gpurun_out/bench_c2_${TAG}.err
cut -c1-400 gpurun_out/bench_c2_${TAG}.json
python bench.py --steps 5 --warmup 3 --batch 1 --arch med3ddram50 --dims 400,512,512 --no-cpu-baseline --no-yardstick --no-train-field > gpurun_out/bench_c4_${TAG}.json 2> gpurun_out/bench_c4_${TAG}.err
cut -c1-400 gpurun_out/bench_c4_${TAG}.json
python tools/aux_bench.py 256 1 > gpurun_out/auxbench_b1_${TAG}.log 2>&1
bash tools/gpu_profile_pipe.sh ${TAG}
bash tools/gpu_profile_pipe.sh ${TAG}_c4 med3ddram50 400,512,512 1
