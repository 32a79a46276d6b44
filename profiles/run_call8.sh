#!/bin/bash
# GPU call 8: LayerNorm variant 2 (precombined affine in registers), cross-Q on single-CTA tiles.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fold_gpu.py -q -s -k "ln_variant2 or cq_single or ksub2" > gpurun_out/c8_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c8_pytest.log
tail -4 gpurun_out/c8_pytest.log
for o in "" "--opt ln_variant=2" "--opt cq_single=1" "--opt ln_variant=2 --opt cq_single=1" "" "--opt ln_variant=2"; do
  timeout 300 python profiles/profile_step.py --steps 1 --vae 0 $o 2>&1 | grep "ms per" >> gpurun_out/c8_ab.txt
done
cat gpurun_out/c8_ab.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c8_launches_warm.csv python profiles/profile_step.py --steps 1 --opt ln_variant=2 --opt cq_single=1 > gpurun_out/c8_ncu1.log 2>&1
EZB_LN_VARIANT=2 timeout 900 python bench.py --no-cpu-baseline --no-extras > gpurun_out/c8_bench_ln2.json 2> gpurun_out/c8_bench.err; cut -c1-300 gpurun_out/c8_bench_ln2.json
ls -la gpurun_out | grep c8_
