#!/bin/bash
# compute-sanitizer passes over the kernel-level GPU tests on small shapes: memcheck on everything listed,
# racecheck / synccheck on the shared-memory-heavy memory-bound kernels.  Slow (10-50x): run on its own gpurun call,
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh TAG'
# and copy gpurun_out/sanitize_TAG.txt into profiles/.  Not run in round 1 (the GPU budget went to parity and timing).
TAG=${1:-r2}
OUT=gpurun_out/sanitize_${TAG}.txt
mkdir -p gpurun_out
: > $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool, pytest selection...
  local tool=$1; shift
  echo "== $tool :: $*" | tee -a $OUT
  timeout 600 $SAN --tool $tool --error-exitcode 9 --launch-timeout 120 \
      python -m pytest "$@" -m gpu -x -q -p no:cacheprovider 2>&1 | tail -6 | tee -a $OUT
  echo "exit code ${PIPESTATUS[0]}" | tee -a $OUT
}
run memcheck tests/test_aux_gpu.py
run memcheck tests/test_prepost_gpu.py
run memcheck tests/test_backward_gpu.py -k "train_loss or adam_kernel or peer_allreduce"
run memcheck tests/test_conv3d_gpu.py -k "not fullsize"
run racecheck tests/test_aux_gpu.py
run synccheck tests/test_backward_gpu.py -k "train_loss or peer_allreduce"
