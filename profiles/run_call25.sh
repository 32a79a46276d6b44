#!/bin/bash
# GPU call 25: attention generation 7 (64-key blocks, two S buffers per group, stores from the Q warp)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_gpu.py -q -x -k "gen7" > gpurun_out/c25_pytest.log 2>&1; RC=$?; echo "pytest exit $RC" >> gpurun_out/c25_pytest.log
tail -12 gpurun_out/c25_pytest.log
if [ $RC -eq 0 ]; then
timeout 200 python profiles/attn_bench.py 0 5 0 > gpurun_out/c25_attn_bench.txt 2>&1; timeout 200 python profiles/attn_bench.py 0 5 1 >> gpurun_out/c25_attn_bench.txt 2>&1; cat gpurun_out/c25_attn_bench.txt
timeout 400 python profiles/ab_sweep.py "" "attn7=1" "" "attn7=1" > gpurun_out/c25_ab.txt 2> gpurun_out/c25_ab.err; cat gpurun_out/c25_ab.txt; tail -3 gpurun_out/c25_ab.err
fi
