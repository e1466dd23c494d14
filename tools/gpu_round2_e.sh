#!/bin/bash
# A/B of the MMA issue order inside a plane-ring weight stage (same sources, -DDRAM_SLAB_MMA_ORDER=0 vs 1).
TAG=${1:-r2e}
mkdir -p gpurun_out
python -m pytest tests/test_conv3d_gpu.py tests/test_training_gpu.py -m gpu -q -x -k "plane_ring or commuted or upconv or lightning" 2>&1 | tail -4
for lib in libdram_b200_order0.so libdram_b200.so; do
  echo "== $lib"
  DRAM_B200_LIB=$PWD/bodyct-dram-emph-subtype_b200/$lib python tools/conv_layer_bench.py 256 1 auto "layer1,us1,us2,us3" 2>&1 | tee gpurun_out/convbench_${lib}_${TAG}.log
done
python tools/engine_profile.py med3ddram 256,256,256 4 > gpurun_out/engine_b4_${TAG}.log 2>&1
grep -v "layer3\.[1-5]\|layer4\.[12]\|layer2\.[123]\|layer1\.[12]" gpurun_out/engine_b4_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cat gpurun_out/bench_b4_${TAG}.json | cut -c1-400; tail -3 gpurun_out/bench_b4_${TAG}.err
