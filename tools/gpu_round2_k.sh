#!/bin/bash
# Is the training step reproducible?  Current library vs the round-2g build (in-order MMA issue, before Cout-128 planes
# and the residual prefetch), three repetitions each; then with the Cout-128 plane-ring kernel switched off.
TAG=${1:-r2k}
mkdir -p gpurun_out
echo "== current library"; python tools/train_repro_check.py 3 2>&1 | tail -12 | tee gpurun_out/train_repro_current_${TAG}.log
echo "== current library, DRAM_B200_SLAB128=0"; DRAM_B200_SLAB128=0 python tools/train_repro_check.py 3 2>&1 | tail -12 | tee gpurun_out/train_repro_noslab128_${TAG}.log
echo "== r2g order0 build"; DRAM_B200_LIB=$PWD/bodyct-dram-emph-subtype_b200/libdram_b200_r2g_order0.so python tools/train_repro_check.py 3 2>&1 | tail -12 | tee gpurun_out/train_repro_r2g_${TAG}.log
for lib in libdram_b200_r2g_order0.so libdram_b200.so; do
  echo "== route test with $lib"
  DRAM_B200_LIB=$PWD/bodyct-dram-emph-subtype_b200/$lib python -m pytest tests/test_training_gpu.py -m gpu -q -s -k "native_loss" 2>&1 | grep "native vs\|passed\|failed"
done
