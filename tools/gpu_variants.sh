#!/bin/bash
# A/B of kernel-tuning builds (build.build_variant) on the plane-ring layers.
#   gpurun --timeout 600 -- 'bash tools/gpu_variants.sh'
mkdir -p gpurun_out
for lib in "" build_variants/*.so; do
  echo "=== ${lib:-default}"
  DRAM_B200_LIB=${lib:+$PWD/$lib} python tools/conv_layer_bench.py 256 ${BATCH:-1} auto "layer1,us1,us2,us3" 2>&1 | grep -v "^TOTAL"
done | tee gpurun_out/variants.log
