#!/bin/bash
# GPU call 11: register-tiled fp32 attention (tests: attention impl 0, bf16x3 DiT, T5), QKV epilogue attribution, T5-XL timing, bench.
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -k "not multicast" > gpurun_out/c11_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c11_pytest.log
tail -5 gpurun_out/c11_pytest.log
timeout 900 bash profiles/heads_dbg.sh > gpurun_out/c11_heads_dbg.txt 2>&1; cat gpurun_out/c11_heads_dbg.txt
timeout 600 python profiles/bench_configs.py --configs T5 > gpurun_out/c11_t5.json 2> gpurun_out/c11_t5.err; cat gpurun_out/c11_t5.json
timeout 900 python bench.py > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; cut -c1-300 gpurun_out/c11_bench.json
ls -la gpurun_out | grep c11_
