#!/bin/bash
# GPU call 21: L2 prefetch of the next GEMM's weights (option w_prefetch)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dit_gpu.py tests/test_gemm_gpu.py -q -x > gpurun_out/c21_pytest.log 2>&1; RC=$?; echo "pytest exit $RC" >> gpurun_out/c21_pytest.log
tail -4 gpurun_out/c21_pytest.log
if [ $RC -eq 0 ]; then
timeout 600 python profiles/ab_sweep.py "" "w_prefetch=0" "" "w_prefetch=0" "" "w_prefetch=0" "skip=16" "skip=16,w_prefetch=0" "skip=8" "skip=8,w_prefetch=0" "skip=4" "skip=4,w_prefetch=0" > gpurun_out/c21_ab.txt 2> gpurun_out/c21_ab.err; cat gpurun_out/c21_ab.txt; tail -3 gpurun_out/c21_ab.err
fi
ls -la gpurun_out | grep c21_
