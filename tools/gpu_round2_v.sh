#!/bin/bash
# heat-map unroll: parity + timing; then one full ncu capture of the stem kernel (only ncu in this call)
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prepost_gpu.py tests/test_guard_gpu.py tests/test_pipeline_gpu.py -m gpu -q -rf > gpurun_out/pytest_${TAG}.log 2>&1
tail -3 gpurun_out/pytest_${TAG}.log
python tools/aux_bench.py 256 1 2>&1 | grep "f1\|f2\|K2 stem" | tee gpurun_out/auxbench_f_${TAG}.log
python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_b4_${TAG}.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_product',round(d['e2e_product']['value'],1))
PY
AUX="python tools/aux_bench.py 256 1"
ncu --set full --clock-control none --import-source on -k regex:'conv3d_stem_kernel' -s 4 -c 1 -o /tmp/stem_${TAG} $AUX > gpurun_out/ncu_stem_${TAG}.log 2>&1
ncu -i /tmp/stem_${TAG}.ncu-rep --page raw --csv > gpurun_out/stem_${TAG}.raw.csv
ncu -i /tmp/stem_${TAG}.ncu-rep --page source --csv > gpurun_out/stem_${TAG}.source.csv 2>/dev/null
ncu -i /tmp/stem_${TAG}.ncu-rep --page details > gpurun_out/stem_${TAG}.details.txt
ls -la gpurun_out/stem_${TAG}.*
