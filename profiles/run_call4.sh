#!/bin/bash
# GPU call 4 of round 2: LayerNorm-tail kernels (never run before: bounded first), attention cycle counters, A/B, bench with the tail on.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_fold_gpu.py -q -s -k "ln_tail" > gpurun_out/c4_pytest_tail.log 2>&1; TRC=$?; echo "tail pytest exit $TRC" >> gpurun_out/c4_pytest_tail.log
tail -5 gpurun_out/c4_pytest_tail.log
KEXPR="not multicast"
if [ $TRC -eq 124 ]; then KEXPR="$KEXPR and not ln_tail"; fi
timeout 1800 python -m pytest tests -m gpu -q -s -k "$KEXPR" > gpurun_out/c4_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c4_pytest.log
tail -8 gpurun_out/c4_pytest.log
timeout 300 python profiles/attn_dbg.py > gpurun_out/c4_attn_dbg.txt 2>&1; cat gpurun_out/c4_attn_dbg.txt
for o in "" "--opt ln_tail=1" "--opt ln_tail=1 --opt dhp80=1" "--opt ln_tail=1 --opt mlp_fused=1"; do
  timeout 300 python profiles/profile_step.py --steps 1 --vae 0 $o 2>&1 | grep "ms per" >> gpurun_out/c4_ab.txt
done
cat gpurun_out/c4_ab.txt
if [ $TRC -ne 124 ]; then
EZB_LN_TAIL=1 EZB_DHP80=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/c4_bench_tail.json 2> gpurun_out/c4_bench_tail.err; echo "bench exit $?"
cut -c1-300 gpurun_out/c4_bench_tail.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c4_launches_warm_tail.csv python profiles/profile_step.py --steps 1 --opt ln_tail=1 > gpurun_out/c4_ncu1.log 2>&1
fi
ls -la gpurun_out | grep c4_
