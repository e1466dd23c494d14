#!/bin/bash
# A/B of kernel-tuning builds (build.build_variant): runs the given command once per library.
#   gpurun --timeout 600 -- 'bash tools/gpu_variants.sh python tools/conv_one.py ...'
mkdir -p gpurun_out
for lib in "" build_variants/*.so; do
  echo "=== ${lib:-default}"
  DRAM_B200_LIB=${lib:+$PWD/$lib} "$@" 2>&1 | tail -8
done | tee gpurun_out/variants.log
