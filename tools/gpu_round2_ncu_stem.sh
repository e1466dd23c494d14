#!/bin/bash
# one full ncu capture of the stem kernel (plain run first)
TAG=${1:-r2aa}
mkdir -p gpurun_out
AUX="python tools/aux_bench.py 256 1"
$AUX 2>&1 | grep "K2 stem"
ncu --set full --clock-control none --import-source on -k regex:'conv3d_stem_kernel' -s 4 -c 1 -o /tmp/stem_${TAG} $AUX > gpurun_out/ncu_stem_${TAG}.log 2>&1
ncu -i /tmp/stem_${TAG}.ncu-rep --page raw --csv > gpurun_out/stem_${TAG}.raw.csv
ncu -i /tmp/stem_${TAG}.ncu-rep --page source --csv > gpurun_out/stem_${TAG}.source.csv 2>/dev/null
ncu -i /tmp/stem_${TAG}.ncu-rep --page details > gpurun_out/stem_${TAG}.details.txt
ls -la gpurun_out/stem_${TAG}.*
