#!/bin/bash
# the driver's own 2-GPU command line (all legs of the default line) on the final tree
TAG=${1:-r2n2e}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_n2_default_${TAG}.json 2> gpurun_out/bench_n2_default_${TAG}.err
echo "rc=$?"; wc -l gpurun_out/bench_n2_default_${TAG}.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2_default_${TAG}.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e_product',round(d['e2e_product']['value'],1), 'keys', sorted(d.keys()))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2_ref_${TAG}.json 2> gpurun_out/bench_n2_ref_${TAG}.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_n2_ref_${TAG}.json
