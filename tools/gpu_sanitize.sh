#!/bin/bash
# ONE compute-sanitizer tool per gpurun call (B200_PROFILING.md: several tools in one call have left a GPU unusable) over
# the kernel-level GPU tests on small shapes.  Slow (10-50x).
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh memcheck|racecheck|synccheck|initcheck TAG'
# and copy gpurun_out/sanitize_<tool>_TAG.txt into profiles/.
TOOL=${1:-memcheck}
TAG=${2:-r2}
OUT=gpurun_out/sanitize_${TOOL}_${TAG}.txt
mkdir -p gpurun_out
: > $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
export DRAM_B200_GRAPH=0   # plain launches: the sanitizer attributes errors per kernel launch
run() {  # pytest selection...
  echo "== $TOOL :: $*" | tee -a $OUT
  timeout 900 $SAN --tool $TOOL --error-exitcode 9 --launch-timeout 120 \
      python -m pytest "$@" -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 | tee -a $OUT
  echo "exit code ${PIPESTATUS[0]}" | tee -a $OUT
}
case $TOOL in
  memcheck)
    run tests/test_aux_gpu.py -k "not opcheck"
    run tests/test_prepost_gpu.py
    run tests/test_conv3d_gpu.py
    run tests/test_backward_gpu.py -k "train_loss or adam_kernel or peer_allreduce or wgrad"
    run tests/test_model_gpu.py -k "oracle or saturation" ;;
  racecheck)  # shared-memory hazards: the mbarrier/TMA pipelines of the conv kernels, K7's staged brick, K4
    run tests/test_conv3d_gpu.py
    run tests/test_aux_gpu.py -k "dram or upsample2x or stem or maxpool" ;;
  synccheck)
    run tests/test_conv3d_gpu.py
    run tests/test_aux_gpu.py -k "dram or upsample2x or stem"
    run tests/test_backward_gpu.py -k "train_loss or peer_allreduce or wgrad" ;;
  initcheck)
    run tests/test_aux_gpu.py -k "not opcheck"
    run tests/test_conv3d_gpu.py ;;
esac
