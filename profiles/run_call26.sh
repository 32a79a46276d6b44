#!/bin/bash
# GPU call 26: final sanity of the committed build (attention kernels of all generations, smoke, XL parity)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -q > gpurun_out/c26_pytest_attn.log 2>&1; echo "pytest exit $?" >> gpurun_out/c26_pytest_attn.log; tail -3 gpurun_out/c26_pytest_attn.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c26_smoke.log 2>&1; tail -5 gpurun_out/c26_smoke.log
timeout 300 python -m pytest tests/test_dit_gpu.py tests/test_api_gpu.py -q -x > gpurun_out/c26_pytest_dit.log 2>&1; echo "pytest exit $?" >> gpurun_out/c26_pytest_dit.log; tail -3 gpurun_out/c26_pytest_dit.log
