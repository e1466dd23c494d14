#!/bin/bash
TAG=${1:-r2n2d}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_n2_${TAG}.json 2> gpurun_out/bench_n2_${TAG}.err
python bench.py --steps 30 --warmup 3 --no-yardstick --no-cpu-baseline --no-train-field > gpurun_out/bench_n1_${TAG}.json 2> gpurun_out/bench_n1_${TAG}.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_n2_${TAG}.json 2> gpurun_out/bench_train_n2_${TAG}.err
python - <<PY
import json
for f in ('gpurun_out/bench_n2_${TAG}.json','gpurun_out/bench_n1_${TAG}.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_product',round(d['e2e_product']['value'],1))
d=json.loads(open('gpurun_out/bench_train_n2_${TAG}.json').read().strip().splitlines()[-1]); print('train n2', round(d['value'],1), d['ms_per_step'])
PY
