#!/bin/bash
TAG=${1:-r2r}
mkdir -p gpurun_out
for n in 1 4; do
  echo "stream+heads n=$n: $(HEADS=1 timeout 120 python tools/conv_one.py $n 128 128 128 64 0 32 3 1 1 | tail -1)"
  echo "ring+heads   n=$n: $(HEADS=1 DRAM_B200_US3=ring timeout 120 python tools/conv_one.py $n 128 128 128 64 0 32 3 1 1 | tail -1)"
done 2>&1 | tee gpurun_out/us3_heads_${TAG}.log
python tools/engine_profile.py med3ddram 256,256,256 4 > gpurun_out/engine_layers_b4_${TAG}.log 2>&1; grep "us3\|engine step" gpurun_out/engine_layers_b4_${TAG}.log
DRAM_B200_US3=ring python tools/engine_profile.py med3ddram 256,256,256 4 2>&1 | grep "us3\|engine step"
python tools/engine_profile.py med3ddram 256,256,256 1 > gpurun_out/engine_layers_b1_${TAG}.log 2>&1; grep "us3\|engine step" gpurun_out/engine_layers_b1_${TAG}.log
