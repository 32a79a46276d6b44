#!/bin/bash
# GPU call 2 of round 2: full suite (incl. folded-LayerNorm tests), fold A/B in situ, GEMM ncu captures, bench with the fold on.
set -x
mkdir -p gpurun_out
# the new attention kernel first, bounded: a dead-lock there must not eat the call
timeout 400 python -m pytest tests/test_attention_gpu.py -q -s > gpurun_out/c2_pytest_attn.log 2>&1; ARC=$?; echo "attention pytest exit $ARC" >> gpurun_out/c2_pytest_attn.log
tail -5 gpurun_out/c2_pytest_attn.log
KEXPR="not multicast and not test_attention"
A5="--opt attn5=1"
if [ $ARC -ne 0 ]; then KEXPR="$KEXPR and not attn5 and not all"; A5=""; fi
timeout 1800 python -m pytest tests -m gpu -q -s -k "$KEXPR" > gpurun_out/c2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c2_pytest.log
tail -15 gpurun_out/c2_pytest.log
timeout 300 python profiles/attn_bench.py > gpurun_out/c2_attn_bench.txt 2>&1; cat gpurun_out/c2_attn_bench.txt
for o in "" "--opt ln_fold=1" "$A5" "--opt dhp80=1" "--opt heads_direct=1" "--opt ln_fold=1 $A5 --opt dhp80=1 --opt heads_direct=1"; do
  timeout 300 python profiles/profile_step.py --steps 1 --vae 0 $o 2>&1 | grep "ms per" >> gpurun_out/c2_ab.txt
done
cat gpurun_out/c2_ab.txt
if [ $ARC -eq 0 ]; then export EZB_ATTN5=1; fi
EZB_LN_FOLD=1 EZB_DHP80=1 EZB_HEADS_DIRECT=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/c2_bench_fold.json 2> gpurun_out/c2_bench_fold.err; echo "bench exit $?"
cut -c1-400 gpurun_out/c2_bench_fold.json
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:gemm -c 9 -o gpurun_out/c2_full_gemm python profiles/profile_step.py --steps 1 --vae 0 --opt ln_fold=1 --opt heads_direct=1 > gpurun_out/c2_ncu1.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c2_launches_warm_fold.csv python profiles/profile_step.py --steps 1 --opt ln_fold=1 > gpurun_out/c2_ncu2.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:gemm_tcgen05_kernel -c 27 -o gpurun_out/c2_full_vae python profiles/profile_step.py --steps 0 --vae 1 > gpurun_out/c2_ncu3.log 2>&1
for r in c2_full_gemm c2_full_vae; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out | grep c2_
