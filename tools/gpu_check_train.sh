#!/bin/bash
# One round-trip on the GPU box for the training tail (K11 loss, K12 Adam, direct wgrad sink): the whole GPU suite,
# smoke, the default bench line, A/B of the training step (native vs ATen loss/optimiser), the training bench line.
#   gpurun --timeout 900 -- 'bash tools/gpu_check_train.sh TAG'
TAG=${1:-r1l}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_${TAG}.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke_${TAG}.log
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_b4_${TAG}.json 2> gpurun_out/bench_b4_${TAG}.err
cat gpurun_out/bench_b4_${TAG}.json
timeout 120 python tools/train_step_bench.py 256 1 resnet34segreg native native > gpurun_out/trainbench_native_${TAG}.log 2>&1
head -5 gpurun_out/trainbench_native_${TAG}.log
timeout 120 python tools/train_step_bench.py 256 1 resnet34segreg aten torch > gpurun_out/trainbench_aten_${TAG}.log 2>&1
head -5 gpurun_out/trainbench_aten_${TAG}.log
timeout 150 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/bench_train_n1_${TAG}.json 2> gpurun_out/bench_train_n1_${TAG}.err
cat gpurun_out/bench_train_n1_${TAG}.json
