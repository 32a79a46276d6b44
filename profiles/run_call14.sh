#!/bin/bash
# GPU call 14: tcgen05.mma / TMA issue under elect.sync in warp-uniform loops (all GEMM kernels + attention): unit tests, attention and GEMM
# micro-benchmarks, one-process option sweep (incl. the retired variants, whose verdicts were measured with the serialised issue loops).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_attention_gpu.py -q -x -k "not multicast" > gpurun_out/c14_pytest_unit.log 2>&1; RC=$?; echo "unit pytest exit $RC" >> gpurun_out/c14_pytest_unit.log
tail -4 gpurun_out/c14_pytest_unit.log
if [ $RC -eq 0 ]; then
for m in 0 2 4 6 1; do timeout 200 python profiles/attn_bench.py $m >> gpurun_out/c14_attn_bench.txt 2>&1; done; cat gpurun_out/c14_attn_bench.txt
timeout 900 python profiles/ab_sweep.py > gpurun_out/c14_ab.txt 2> gpurun_out/c14_ab.err; cat gpurun_out/c14_ab.txt; tail -3 gpurun_out/c14_ab.err
timeout 300 python profiles/gemm_bench.py > gpurun_out/c14_gemm_bench.txt 2>&1; cat gpurun_out/c14_gemm_bench.txt
timeout 900 python -m pytest tests/test_dit_gpu.py tests/test_loop_gpu.py tests/test_vae_gpu.py -q -s > gpurun_out/c14_pytest_parity.log 2>&1; echo "parity pytest exit $?" >> gpurun_out/c14_pytest_parity.log
tail -4 gpurun_out/c14_pytest_parity.log
fi
ls -la gpurun_out | grep c14_
