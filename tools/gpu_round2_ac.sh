#!/bin/bash
TAG=${1:-r2ac}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py tests/test_conv3d_gpu.py -m gpu -q -rf -k "not properties" > gpurun_out/pytest_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_b4_${TAG}.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_product',round(d['e2e_product']['value'],1))
PY
