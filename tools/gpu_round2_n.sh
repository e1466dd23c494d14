#!/bin/bash
# us3 streaming kernel: parity, then timing against the 4-plane items, then the whole step.
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv3d_gpu.py -m gpu -q -rf -x -k "streaming or plane_ring or reproducible or dispatch" > gpurun_out/pytest_${TAG}.log 2>&1
tail -6 gpurun_out/pytest_${TAG}.log
if ! grep -q " passed" gpurun_out/pytest_${TAG}.log || grep -q "failed\|error" gpurun_out/pytest_${TAG}.log; then echo "PARITY NOT GREEN - stopping"; exit 1; fi
for n in 1 4; do
  echo "stream  n=$n: $(timeout 120 python tools/conv_one.py $n 128 128 128 64 0 32 3 1 1 | tail -1)"
  echo "ring    n=$n: $(DRAM_B200_US3=ring timeout 120 python tools/conv_one.py $n 128 128 128 64 0 32 3 1 1 | tail -1)"
done 2>&1 | tee gpurun_out/us3_stream_${TAG}.log
[ "$2" = "quick" ] && exit 0
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py -m gpu -q -rf -k "not properties" > gpurun_out/pytest_model_${TAG}.log 2>&1
tail -4 gpurun_out/pytest_model_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cut -c1-330 gpurun_out/bench_b4_${TAG}.json; tail -3 gpurun_out/bench_b4_${TAG}.err
python bench.py --steps 40 --warmup 3 --batch 1 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_b1_${TAG}.json 2> gpurun_out/bench_b1_${TAG}.err
cut -c1-330 gpurun_out/bench_b1_${TAG}.json
