#!/bin/bash
# GPU call 1 of round 2: full GPU suite, microbenchmarks, in-situ A/B timings, bench line, launch list + ncu --set full captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s -k "not multicast" > gpurun_out/c1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 120 profiles/micro/microbench > gpurun_out/c1_microbench.txt 2>&1
# in-situ cost of each kernel class (skip mask) and the multicast switch
for o in "" "--opt skip=1" "--opt skip=2" "--opt skip=4" "--opt skip=8" "--opt skip=16"; do
  timeout 300 python profiles/profile_step.py --steps 1 --vae 0 $o 2>&1 | grep "ms per" >> gpurun_out/c1_ab.txt
done
cat gpurun_out/c1_ab.txt
timeout 900 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench exit $?"
cat gpurun_out/c1_bench.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c1_bench_ref.json 2> gpurun_out/c1_bench_ref.err
# launch list (warm caches: no flush between kernels) of one step + one decode
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c1_launches_warm.csv python profiles/profile_step.py --steps 1 > gpurun_out/c1_ncu1.log 2>&1
# ncu --set full for the kernels that had no capture (reports converted to CSV here; the .ncu-rep files are dropped: gpurun_out is capped at 64 MiB)
timeout 1200 ncu --profile-from-start off --set full --clock-control none \
  -k regex:"ln_mod_cast_reg|EpiHeads|attn4|EpiLinearT|EpiGeglu" -c 11 -o gpurun_out/c1_full_block python profiles/profile_step.py --steps 1 --vae 0 > gpurun_out/c1_ncu3.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none \
  -k regex:"cfg_ddim|final_conv|wave_out|patch_pack|latent_pack" -c 5 -o gpurun_out/c1_full_misc python profiles/profile_step.py --steps 1 --vae 1 > gpurun_out/c1_ncu4.log 2>&1
timeout 1200 ncu --profile-from-start off --set full --clock-control none \
  -k regex:"gemm_tcgen05_kernel<128" -c 27 -o gpurun_out/c1_full_vae python profiles/profile_step.py --steps 0 --vae 1 > gpurun_out/c1_ncu5.log 2>&1
for r in c1_full_block c1_full_misc c1_full_vae; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
# last (never run on hardware before): the multicast swap-AB GEMM
timeout 120 python -m pytest tests/test_gemm_gpu.py -x -q -k multicast > gpurun_out/c1_mc_test.log 2>&1; echo "mc test exit $?" >> gpurun_out/c1_mc_test.log
tail -3 gpurun_out/c1_mc_test.log
timeout 200 python profiles/profile_step.py --steps 1 --vae 0 --opt swap_mc=1 2>&1 | grep "ms per" >> gpurun_out/c1_ab.txt
tail -2 gpurun_out/c1_ab.txt
ls -la gpurun_out | tail -20
