#!/bin/bash
# training-route distance with both MMA orders, then compute-sanitizer memcheck (the only tool of this call)
TAG=${1:-r2g}
mkdir -p gpurun_out
for lib in libdram_b200_order0.so libdram_b200.so; do
  echo "== $lib"
  DRAM_B200_LIB=$PWD/bodyct-dram-emph-subtype_b200/$lib python -m pytest tests/test_training_gpu.py -m gpu -q -s -k "native_loss" 2>&1 | grep "weight gradient\|native vs\|passed\|failed" | tee -a gpurun_out/train_route_${TAG}.log
done
bash tools/gpu_sanitize.sh memcheck ${TAG}
