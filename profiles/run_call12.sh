#!/bin/bash
# GPU call 12 (2 GPUs): the torchrun launch the driver uses for the scaling run, N = 2 (ours + reference arm).
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c12_bench_n2.json 2> gpurun_out/c12_bench_n2.err; echo "n2 exit $?"
cut -c1-400 gpurun_out/c12_bench_n2.json; tail -3 gpurun_out/c12_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/c12_bench_ref_n2.json 2> gpurun_out/c12_bench_ref_n2.err; echo "ref exit $?"
cut -c1-200 gpurun_out/c12_bench_ref_n2.json
