#!/bin/bash
# GPU call 6: whole-schedule CUDA graph + dhp80 default: loop / API / fold tests, bench.
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -k "not multicast" > gpurun_out/c6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c6_pytest.log
tail -6 gpurun_out/c6_pytest.log
timeout 900 python bench.py > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; echo "bench exit $?"
cut -c1-300 gpurun_out/c6_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c6_bench_ref.json 2> gpurun_out/c6_bench_ref.err
ls -la gpurun_out | grep c6_
