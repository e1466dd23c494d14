#!/bin/bash
# full ncu capture of the staged 1x1x1 kernel on ResNet-50's layer1 conv3 shape at C4 size (64 -> 256 + residual)
TAG=${1:-r2k1}
mkdir -p gpurun_out
C3="python tools/conv_one.py 1 100 128 128 64 0 256 1 1 1 256"
C1="python tools/conv_one.py 1 100 128 128 256 0 64 1 1 1"
$C3 | tail -1; $C1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:'conv3d_umma_kernel' -s 3 -c 1 -o /tmp/c3_${TAG} $C3 > gpurun_out/ncu_c3_${TAG}.log 2>&1
ncu -i /tmp/c3_${TAG}.ncu-rep --page raw --csv > gpurun_out/c3_${TAG}.raw.csv
ncu -i /tmp/c3_${TAG}.ncu-rep --page source --csv > gpurun_out/c3_${TAG}.source.csv 2>/dev/null
ncu -i /tmp/c3_${TAG}.ncu-rep --page details > gpurun_out/c3_${TAG}.details.txt
ls -la gpurun_out/c3_${TAG}.*
