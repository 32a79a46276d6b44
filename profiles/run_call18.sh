#!/bin/bash
# GPU call 18: attention generation 6 (chunked two-pass softmax, packed arithmetic, MUFU token, P in halves)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -q -x -k "gen6" > gpurun_out/c18_pytest_attn.log 2>&1; RC=$?; echo "pytest exit $RC" >> gpurun_out/c18_pytest_attn.log
tail -15 gpurun_out/c18_pytest_attn.log
for m in 0 1 3 5 7; do timeout 200 python profiles/attn_bench.py 0 $m 2>&1 | grep -v "impl   5" >> gpurun_out/c18_attn_bench.txt; done; cat gpurun_out/c18_attn_bench.txt
if [ $RC -eq 0 ]; then
timeout 600 python profiles/ab_sweep.py "" "attn6=1" "attn6=3" "attn6=5" "attn6=7" "" "attn6=7" > gpurun_out/c18_ab.txt 2> gpurun_out/c18_ab.err; cat gpurun_out/c18_ab.txt; tail -3 gpurun_out/c18_ab.err
fi
ls -la gpurun_out | grep c18_
