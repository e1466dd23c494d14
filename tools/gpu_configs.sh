#!/bin/bash
# BASELINE.json configurations as bench lines (C2, C3 at batch 1 and 4, C4) + the full-size tests.
#   gpurun --timeout 1200 -- 'bash tools/gpu_configs.sh TAG'
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_fullsize_${TAG}.log
python bench.py --arch med3ddram18 --batch 4 --steps 10 --no-cpu-baseline > gpurun_out/bench_c2_${TAG}.json 2> gpurun_out/bench_c2_${TAG}.err
python bench.py --arch med3ddram50 --dims 400,512,512 --batch 1 --steps 5 --no-cpu-baseline > gpurun_out/bench_c4_${TAG}.json 2> gpurun_out/bench_c4_${TAG}.err
tail -c 600 gpurun_out/bench_c4_${TAG}.err
python - <<PY
import json
for c in ("c2", "c4"):
    try:
        d = json.load(open(f"gpurun_out/bench_{c}_${TAG}.json"))
        print(c, d["config"]["workload"][:60], "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2),
              "e2e", round(d["e2e"]["value"], 2), "frac", round(d["roofline"]["frac"], 3), d["clocks"])
    except Exception as e:
        print(c, "failed", e)
PY
