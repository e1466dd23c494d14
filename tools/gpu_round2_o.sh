#!/bin/bash
# Full ncu capture of the us3 streaming kernel and a 64-channel plane-ring launch for comparison (one GPU, one tool).
TAG=${1:-r2o}
mkdir -p gpurun_out
US3="python tools/conv_one.py 1 128 128 128 64 0 32 3 1 1"
L64="python tools/conv_one.py 1 128 128 128 64 0 64 3 1 1"
$US3 | tail -1; $L64 | tail -1
ncu --set full --clock-control none --import-source on -k regex:'conv3d_stream32' -s 3 -c 1 -o /tmp/us3_${TAG} $US3 > gpurun_out/ncu_us3_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv3d_slab' -s 3 -c 1 -o /tmp/l64_${TAG} $L64 > gpurun_out/ncu_l64_${TAG}.log 2>&1
for k in us3 l64; do
  ncu -i /tmp/${k}_${TAG}.ncu-rep --page raw --csv > gpurun_out/${k}_${TAG}.raw.csv
  ncu -i /tmp/${k}_${TAG}.ncu-rep --page source --csv > gpurun_out/${k}_${TAG}.source.csv 2>/dev/null
  ncu -i /tmp/${k}_${TAG}.ncu-rep --page details > gpurun_out/${k}_${TAG}.details.txt
done
ls -la gpurun_out/*_${TAG}.*
