#!/bin/bash
# GPU call 7: swap-AB epilogue with pipelined residual prefetch (default path), 128-deep stages for the wide pair GEMMs (option ksub2).
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -k "not multicast" > gpurun_out/c7_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c7_pytest.log
tail -5 gpurun_out/c7_pytest.log
for o in "" "--opt ksub2=1" "" "--opt ksub2=1"; do
  timeout 300 python profiles/profile_step.py --steps 1 --vae 0 $o 2>&1 | grep "ms per" >> gpurun_out/c7_ab.txt
done
cat gpurun_out/c7_ab.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c7_launches_warm_ksub2.csv python profiles/profile_step.py --steps 1 --opt ksub2=1 > gpurun_out/c7_ncu1.log 2>&1
EZB_KSUB2=1 timeout 900 python bench.py --no-cpu-baseline --no-extras > gpurun_out/c7_bench_ksub2.json 2> gpurun_out/c7_bench.err; cut -c1-300 gpurun_out/c7_bench_ksub2.json
timeout 900 python bench.py --no-cpu-baseline --no-extras > gpurun_out/c7_bench_default.json 2>> gpurun_out/c7_bench.err; cut -c1-300 gpurun_out/c7_bench_default.json
ls -la gpurun_out | grep c7_
