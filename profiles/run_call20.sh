#!/bin/bash
# GPU call 20: generation 6 as the default attention (P in halves, mask bits fetched before the S wait); gen-5 / two-issuer variants removed
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -q -x > gpurun_out/c20_pytest_attn.log 2>&1; RC=$?; echo "pytest exit $RC" >> gpurun_out/c20_pytest_attn.log
tail -8 gpurun_out/c20_pytest_attn.log
for m in 0 1 5 7; do timeout 200 python profiles/attn_bench.py 0 $m 2>&1 >> gpurun_out/c20_attn_bench.txt; done; cat gpurun_out/c20_attn_bench.txt
if [ $RC -eq 0 ]; then
timeout 900 python -m pytest tests/test_fold_gpu.py -q -x -k "fast_path_options" > gpurun_out/c20_pytest_opts.log 2>&1; echo "pytest exit $?" >> gpurun_out/c20_pytest_opts.log; tail -4 gpurun_out/c20_pytest_opts.log
timeout 600 python profiles/ab_sweep.py "" "attn6=0" "attn6=1" "attn6=7" "" "attn6=0" "skip=2" > gpurun_out/c20_ab.txt 2> gpurun_out/c20_ab.err; cat gpurun_out/c20_ab.txt; tail -3 gpurun_out/c20_ab.err
fi
ls -la gpurun_out | grep c20_
