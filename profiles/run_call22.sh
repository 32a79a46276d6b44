#!/bin/bash
# GPU call 22: validation of the round-2b build (elect-issued tcgen05 / TMA, attention generation 6, wave_out): GPU suite without the 7-minute option
# matrix (validated in call 20), smoke, bench line, warm launch list, ncu --set full captures (GEGLU for roofline.traffic, attention, QKV, swap-AB, wave_out)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -k "not fast_path_options and not multicast" > gpurun_out/c22_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c22_pytest.log
tail -4 gpurun_out/c22_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c22_smoke.log 2>&1; tail -5 gpurun_out/c22_smoke.log
timeout 900 python bench.py > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/c22_bench.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c22_launches_warm.csv python profiles/profile_step.py --steps 1 > gpurun_out/c22_ncu1.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"EpiHeads|attn6|EpiLinearT|EpiGeglu|ln_gc" -c 12 -f -o gpurun_out/c22_full_block python profiles/profile_step.py --steps 1 --vae 0 > gpurun_out/c22_ncu2.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"wave_out|cfg_ddim" -c 2 -f -o gpurun_out/c22_full_misc python profiles/profile_step.py --steps 1 --vae 1 > gpurun_out/c22_ncu3.log 2>&1
for r in c22_full_block c22_full_misc; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out | grep c22_
