#!/bin/bash
# 2 GPUs: which copy leg costs the end-to-end loop its overlap?  (diagnostic A/B, same box)
TAG=${1:-r2n2c}
mkdir -p gpurun_out
for diag in none nod2h noh2d; do
E=""; [ $diag != none ] && E=$diag
DRAM_B200_E2E_DIAG=$E python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_n2_${diag}_${TAG}.json 2> gpurun_out/bench_n2_${diag}_${TAG}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2_${diag}_${TAG}.json').read().strip().splitlines()[-1])
print('${diag}','value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_product',round(d['e2e_product']['value'],1))
PY
done
DRAM_B200_E2E_DIAG=nod2h python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1 nod2h value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))"
DRAM_B200_E2E_DIAG=noh2d python bench.py --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1 noh2d value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))"
