#!/bin/bash
# GPU call 3 of round 2: baseline restored?  LayerNorm kernel variants, attention time attribution, fused-MLP experiment (last: never run before).
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -k "not multicast" > gpurun_out/c3_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c3_pytest.log
tail -6 gpurun_out/c3_pytest.log
timeout 120 profiles/micro/microbench > gpurun_out/c3_microbench.txt 2>&1
timeout 300 python profiles/attn_dbg.py > gpurun_out/c3_attn_dbg.txt 2>&1; cat gpurun_out/c3_attn_dbg.txt
for o in "" "--opt ln_variant=1" "--opt heads_direct=1" "--opt dhp80=1" "--opt ln_fold=1"; do
  timeout 300 python profiles/profile_step.py --steps 1 --vae 0 $o 2>&1 | grep "ms per" >> gpurun_out/c3_ab.txt
done
cat gpurun_out/c3_ab.txt
timeout 900 python bench.py > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench exit $?"
cut -c1-300 gpurun_out/c3_bench.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c3_launches_warm.csv python profiles/profile_step.py --steps 1 > gpurun_out/c3_ncu1.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:gemm -c 9 -o gpurun_out/c3_full_gemm python profiles/profile_step.py --steps 1 --vae 0 > gpurun_out/c3_ncu2.log 2>&1
ncu -i gpurun_out/c3_full_gemm.ncu-rep --page raw --csv > gpurun_out/c3_full_gemm.raw.csv 2>/dev/null; rm -f gpurun_out/c3_full_gemm.ncu-rep
# experimental, never run before: one persistent launch for the MLP (mixes cta_group::2 and ::1 tcgen05 in one kernel)
timeout 150 python profiles/profile_step.py --steps 1 --vae 0 --opt mlp_fused=1 2>&1 | grep -E "ms per|rror" >> gpurun_out/c3_ab.txt; echo "mlp_fused exit $?" >> gpurun_out/c3_ab.txt
tail -3 gpurun_out/c3_ab.txt
ls -la gpurun_out | grep c3_
