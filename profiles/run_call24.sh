#!/bin/bash
# GPU call 24: key-mask bytes requested one block ahead (generation-6 attention)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -q -x -k "gen6 or test_attention[" > gpurun_out/c24_pytest.log 2>&1; RC=$?; echo "pytest exit $RC" >> gpurun_out/c24_pytest.log
tail -4 gpurun_out/c24_pytest.log
timeout 200 python profiles/attn_bench.py 0 5 > gpurun_out/c24_attn_bench.txt 2>&1; cat gpurun_out/c24_attn_bench.txt
timeout 300 python -m pytest tests/test_dit_gpu.py -q -x -k "XL" > gpurun_out/c24_pytest_dit.log 2>&1; echo "pytest exit $?" >> gpurun_out/c24_pytest_dit.log; tail -3 gpurun_out/c24_pytest_dit.log
