#!/bin/bash
# Single-pass ncu evidence for the "conv3d tensor-pipe util %" half of BASELINE's metric (the
# --set full multi-pass replay gave 1.3 / 2.7 / 11.5 % for identical launches).  A short metric list that ncu collects
# in ONE pass per kernel, over the conv launches of the last two of three eager engine passes, + the launch list of a
# bench run.  Only ncu in this call (B200_PROFILING.md), each after its plain run exited 0.
#   gpurun --timeout 1200 -- 'bash tools/gpu_profile_pipe.sh TAG [arch] [D,H,W] [batch]'
TAG=${1:-r2}
ARCH=${2:-med3ddram}
DIMS=${3:-256,256,256}
B=${4:-1}
mkdir -p gpurun_out
CMD="python tools/engine_steps_dump.py gpurun_out/steps_${TAG}.json $ARCH $DIMS $B 3"
METRICS=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__cycles_active.avg
$CMD > gpurun_out/pipe_plain_${TAG}.log 2>&1 &&
ncu --metrics $METRICS --clock-control none -k regex:'conv3d_|upsample2x_umma' --csv \
    --log-file gpurun_out/pipe_${TAG}.csv $CMD > gpurun_out/ncu_pipe_${TAG}.log 2>&1
echo "pipe capture rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --batch $B --arch $ARCH --dims $DIMS --no-cpu-baseline --no-yardstick"
$BENCH > gpurun_out/prof_bench_${TAG}.json 2> gpurun_out/prof_bench_${TAG}.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
# DRAM traffic + issue detail of the memory-bound kernels (also single metric list, one pass)
AUX="python tools/aux_bench.py 256 1"
$AUX > gpurun_out/aux_plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:'dram_upsample_mask|window_|maxpool3d|masked_pool_partial|conv3d_stem' --csv \
    --log-file gpurun_out/auxpipe_${TAG}.csv $AUX > gpurun_out/ncu_aux_${TAG}.log 2>&1
echo "aux capture rc=$?"
ls -la gpurun_out | tail -12
