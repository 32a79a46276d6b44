#!/bin/bash
# GPU call 28: fp32 CUDA-core attention with vector shared-memory loads and probabilities through shared memory (T5, bf16x3 parity mode)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_t5_gpu.py -q -k "test_attention[0- or t5" > gpurun_out/c28_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c28_pytest.log; tail -3 gpurun_out/c28_pytest.log
timeout 200 python profiles/bench_configs.py --configs T5 > gpurun_out/c28_t5.json 2> gpurun_out/c28_t5.err; cat gpurun_out/c28_t5.json | cut -c1-400; tail -2 gpurun_out/c28_t5.err
timeout 300 python -m pytest tests/test_dit_gpu.py -q -k "bf16x3 and (tiny72 or XL)" > gpurun_out/c28_pytest_dit.log 2>&1; echo "pytest exit $?" >> gpurun_out/c28_pytest_dit.log; tail -3 gpurun_out/c28_pytest_dit.log
