#!/bin/bash
# f1 / f2 rework: parity (pre/post, guard bands, pipeline, processor) then timings and the e2e legs.
TAG=${1:-r2u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prepost_gpu.py tests/test_guard_gpu.py tests/test_pipeline_gpu.py tests/test_aux_gpu.py -m gpu -q -rf > gpurun_out/pytest_${TAG}.log 2>&1
tail -6 gpurun_out/pytest_${TAG}.log
python tools/aux_bench.py 256 1 2>&1 | grep "f1\|f2" | tee gpurun_out/auxbench_f_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_b4_${TAG}.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_product',round(d['e2e_product']['value'],1))
PY
tail -2 gpurun_out/bench_b4_${TAG}.err
