#!/bin/bash
# cuDNN yard-stick (the reference's ATen ops on the same B200) for the other BASELINE configurations
TAG=${1:-r2yard}
mkdir -p gpurun_out
python bench.py --impl cudnn --arch med3ddram18 --steps 10 --warmup 3 > gpurun_out/bench_cudnn_c2_${TAG}.json 2> gpurun_out/bench_cudnn_c2_${TAG}.err
cut -c1-400 gpurun_out/bench_cudnn_c2_${TAG}.json
python bench.py --impl cudnn --batch 1 --steps 10 --warmup 3 > gpurun_out/bench_cudnn_b1_${TAG}.json 2> gpurun_out/bench_cudnn_b1_${TAG}.err
cut -c1-400 gpurun_out/bench_cudnn_b1_${TAG}.json
timeout 600 python bench.py --impl cudnn --arch med3ddram50 --dims 400,512,512 --batch 1 --steps 3 --warmup 3 > gpurun_out/bench_cudnn_c4_${TAG}.json 2> gpurun_out/bench_cudnn_c4_${TAG}.err
cut -c1-400 gpurun_out/bench_cudnn_c4_${TAG}.json; tail -2 gpurun_out/bench_cudnn_c4_${TAG}.err
