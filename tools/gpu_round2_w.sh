#!/bin/bash
# stem epilogue rework: parity (stem cases, reproducibility, guard bands, model goldens) then timing
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv3d_gpu.py tests/test_guard_gpu.py tests/test_aux_gpu.py -m gpu -q -rf -x -k "stem or guard or sentinel" > gpurun_out/pytest_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_${TAG}.log
if grep -q "failed\|error" gpurun_out/pytest_${TAG}.log; then echo "PARITY NOT GREEN"; exit 1; fi
python tools/aux_bench.py 256 1 2>&1 | grep "K2 stem" | tee gpurun_out/auxbench_stem_${TAG}.log
python tools/aux_bench.py 256 4 2>&1 | grep "K2 stem" | tee -a gpurun_out/auxbench_stem_${TAG}.log
[ "$2" = "quick" ] && exit 0
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py -m gpu -q -rf -k "not properties" > gpurun_out/pytest_model_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_model_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
python bench.py --steps 40 --warmup 3 --batch 1 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_b1_${TAG}.json 2> gpurun_out/bench_b1_${TAG}.err
python - <<PY
import json
for f in ('gpurun_out/bench_b4_${TAG}.json','gpurun_out/bench_b1_${TAG}.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_product',round(d['e2e_product']['value'],1))
PY
