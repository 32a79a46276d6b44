#!/bin/bash
# GPU call 23: final_conv with four input-channel groups; per-class in-situ attribution of the final build; ncu captures of the block GEMMs and a
# source-level capture of the generation-6 attention kernel; bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dit_gpu.py tests/test_loop_gpu.py -q -x > gpurun_out/c23_pytest.log 2>&1; RC=$?; echo "pytest exit $RC" >> gpurun_out/c23_pytest.log
tail -4 gpurun_out/c23_pytest.log
if [ $RC -eq 0 ]; then
timeout 600 python profiles/ab_sweep.py "" "skip=1" "skip=2" "skip=4" "skip=8" "skip=16" "" > gpurun_out/c23_ab.txt 2> gpurun_out/c23_ab.err; cat gpurun_out/c23_ab.txt; tail -3 gpurun_out/c23_ab.err
timeout 900 python bench.py > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/c23_bench.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c23_launches_warm.csv python profiles/profile_step.py --steps 1 > gpurun_out/c23_ncu1.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:gemm -c 8 -f -o gpurun_out/c23_full_gemms python profiles/profile_step.py --steps 1 --vae 0 > gpurun_out/c23_ncu2.log 2>&1
ncu -i gpurun_out/c23_full_gemms.ncu-rep --page raw --csv > gpurun_out/c23_full_gemms.raw.csv 2>/dev/null; rm -f gpurun_out/c23_full_gemms.ncu-rep
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn6 -s 3 -c 1 -f -o gpurun_out/c23_attn6_self python profiles/attn_one.py self 3 > gpurun_out/c23_ncu_self.log 2>&1; tail -2 gpurun_out/c23_ncu_self.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn6 -s 3 -c 1 -f -o gpurun_out/c23_attn6_cross python profiles/attn_one.py cross 3 > gpurun_out/c23_ncu_cross.log 2>&1; tail -2 gpurun_out/c23_ncu_cross.log
fi
ls -la gpurun_out | grep c23_
