#!/bin/bash
# 8 GPUs, final tree: inference line only (the CPU / cuDNN / training legs run on rank 0 alone and would idle 7 GPUs)
TAG=${1:-r2n8b}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 30 --warmup 3 \
    --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_n8_${TAG}.json 2> gpurun_out/bench_n8_${TAG}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n8_${TAG}.json').read().strip().splitlines()[-1])
print('n8 value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_product',round(d['e2e_product']['value'],1), d['config'].get('cpu_binding'))
PY
tail -2 gpurun_out/bench_n8_${TAG}.err
nvidia-smi topo -m > gpurun_out/topo_${TAG}.txt 2>&1
