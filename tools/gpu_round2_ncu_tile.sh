#!/bin/bash
# full ncu capture of the per-tap tile kernel on the layer3 / layer4 shapes at batch 4 (plain run first)
TAG=${1:-r2tile}
mkdir -p gpurun_out
L3="python tools/conv_one.py 4 32 32 32 256 0 256 3 1 2"
L4="python tools/conv_one.py 4 32 32 32 512 0 512 3 1 4"
$L3 | tail -1; $L4 | tail -1
ncu --set full --clock-control none --import-source on -k regex:'conv3d_umma_kernel' -s 3 -c 1 -o /tmp/l3_${TAG} $L3 > gpurun_out/ncu_l3_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv3d_umma_kernel|conv3d_pair_kernel' -s 3 -c 1 -o /tmp/l4_${TAG} $L4 > gpurun_out/ncu_l4_${TAG}.log 2>&1
for k in l3 l4; do
  ncu -i /tmp/${k}_${TAG}.ncu-rep --page raw --csv > gpurun_out/${k}_${TAG}.raw.csv
  ncu -i /tmp/${k}_${TAG}.ncu-rep --page source --csv > gpurun_out/${k}_${TAG}.source.csv 2>/dev/null
  ncu -i /tmp/${k}_${TAG}.ncu-rep --page details > gpurun_out/${k}_${TAG}.details.txt
done
ls -la gpurun_out/*_${TAG}.*
