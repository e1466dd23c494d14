#!/bin/bash
# per-quarter staged epilogue: parity (1x1x1 staged cases, K13, ResNet-50 goldens, C4 full size) then timing
TAG=${1:-r2k2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv3d_gpu.py tests/test_guard_gpu.py -m gpu -q -rf -x -k "staged or 1x1 or commuted or upconv or guard or sentinel" > gpurun_out/pytest_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_${TAG}.log
if grep -q "failed\|error" gpurun_out/pytest_${TAG}.log; then echo "PARITY NOT GREEN"; exit 1; fi
python tools/conv_one.py 1 100 128 128 64 0 256 1 1 1 256 | tail -1
python tools/conv_one.py 1 100 128 128 256 0 64 1 1 1 | tail -1
python tools/conv_one.py 1 50 64 64 512 0 128 1 1 1 | tail -1
python tools/conv_one.py 1 25 32 32 512 0 2048 1 1 1 2048 | tail -1
[ "$2" = "quick" ] && exit 0
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py -m gpu -q -rf > gpurun_out/pytest_model_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_model_${TAG}.log
python bench.py --steps 5 --warmup 3 --batch 1 --arch med3ddram50 --dims 400,512,512 --no-cpu-baseline --no-yardstick --no-train-field > gpurun_out/bench_c4_${TAG}.json 2> gpurun_out/bench_c4_${TAG}.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-yardstick --no-train-field > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
python - <<PY
import json
for f in ('gpurun_out/bench_c4_${TAG}.json','gpurun_out/bench_b4_${TAG}.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'ms',round(d['ms_per_step'],2))
PY
