#!/bin/bash
# ncu evidence for profiles/: launch list of one bench run + full captures of the conv kernels and the
# memory-bound kernels.  Run only after the same bench command has exited 0 without ncu.  Reports are
# exported to CSV on the box and the .ncu-rep files dropped (gpurun_out/ is capped at 64 MiB).
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh TAG [launches|conv|aux|stem ...]'
TAG=${1:-r1}
shift
WHAT=${@:-launches conv aux}
BENCH="python bench.py --steps 2 --warmup 3 --batch 1 --no-cpu-baseline"
mkdir -p gpurun_out
if [ "$WHAT" != "traintail" ]; then
  $BENCH > gpurun_out/prof_bench_${TAG}.json 2> gpurun_out/prof_bench_${TAG}.err || exit 1
fi
for w in $WHAT; do
  case $w in
    launches)
      ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${TAG}.csv \
          $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1 ;;
    conv)  # one full step of conv launches (the second timed step; 38 conv launches per step)
      ncu --set full --clock-control none -k regex:'conv3d_' -s 152 -c 38 -o /tmp/conv_${TAG} \
          $BENCH > gpurun_out/ncu_conv_${TAG}.log 2>&1
      ncu -i /tmp/conv_${TAG}.ncu-rep --page raw --csv > gpurun_out/conv_${TAG}.raw.csv ;;
    stem)
      ncu --set full --clock-control none --import-source on -k regex:'conv3d_stem' -s 4 -c 1 -o /tmp/stem_${TAG} \
          $BENCH > gpurun_out/ncu_stem_${TAG}.log 2>&1
      ncu -i /tmp/stem_${TAG}.ncu-rep --page raw --csv > gpurun_out/stem_${TAG}.raw.csv
      ncu -i /tmp/stem_${TAG}.ncu-rep --page source --csv > gpurun_out/stem_${TAG}.src.csv
      ncu -i /tmp/stem_${TAG}.ncu-rep --page details > gpurun_out/stem_${TAG}.details.txt ;;
    traintail)  # training step: the loss (K11) and Adam (K12) of the first two steps.  Keep it this small: under
      # --set full every profiled kernel is replayed ~39 times with the step's whole footprint saved and restored
      # (≈ 8 s per kernel at 256^3; a 50-kernel capture did not finish in 7 minutes in round 1).
      ncu --set full --clock-control none -k regex:'adam_kernel|loss_sums|loss_backward|loss_finalize' -c 8 \
          -o /tmp/tt_${TAG} python tools/train_step_bench.py 128 1 resnet34segreg native native > gpurun_out/ncu_traintail_${TAG}.log 2>&1
      ncu -i /tmp/tt_${TAG}.ncu-rep --page raw --csv > gpurun_out/traintail_${TAG}.raw.csv ;;
    aux)
      ncu --set full --clock-control none \
          -k regex:'upsample2x|maxpool3d|dram_upsample_mask|window_|masked_pool_partial' -s 36 -c 9 -o /tmp/aux_${TAG} \
          $BENCH > gpurun_out/ncu_aux_${TAG}.log 2>&1
      ncu -i /tmp/aux_${TAG}.ncu-rep --page raw --csv > gpurun_out/aux_${TAG}.raw.csv ;;
  esac
done
ls -la gpurun_out | tail -12
