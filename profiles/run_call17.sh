#!/bin/bash
# GPU call 17: MUFU-token ping-pong of the two softmax groups (option attn_pp)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -q -x -k "mufu_token" > gpurun_out/c17_pytest_attn.log 2>&1; RC=$?; echo "pytest exit $RC" >> gpurun_out/c17_pytest_attn.log
tail -4 gpurun_out/c17_pytest_attn.log
if [ $RC -eq 0 ]; then
for m in 0 8 10 12 9 14; do timeout 200 python profiles/attn_bench.py $m 2>&1 | grep -v "impl   5" >> gpurun_out/c17_attn_bench.txt; done; cat gpurun_out/c17_attn_bench.txt
timeout 600 python profiles/ab_sweep.py "" "attn_pp=1" "attn_pp=1,attn_poly=1" "attn_pp=1,attn_res=1" "attn_pp=1,attn_mma2=1" "" "attn_pp=1" > gpurun_out/c17_ab.txt 2> gpurun_out/c17_ab.err; cat gpurun_out/c17_ab.txt; tail -3 gpurun_out/c17_ab.err
fi
ls -la gpurun_out | grep c17_
