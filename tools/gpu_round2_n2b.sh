#!/bin/bash
# 2 GPUs: NUMA binding on / off (A/B on the same box), inference line
TAG=${1:-r2n2b}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${TAG}.txt 2>&1
for numa in 1 0; do
DRAM_B200_NUMA=$numa python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_n2_numa${numa}_${TAG}.json 2> gpurun_out/bench_n2_numa${numa}_${TAG}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2_numa${numa}_${TAG}.json').read().strip().splitlines()[-1])
print('numa',${numa},'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_product',round(d['e2e_product']['value'],1), d['config'].get('cpu_binding'))
PY
done
