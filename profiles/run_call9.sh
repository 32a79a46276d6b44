#!/bin/bash
# GPU call 9: defaults = whole-schedule graph + dhp80 + ksub2 (GEGLU) + LayerNorm variant 2 (+ register-resident skip_norm): full suite, bench, launch list.
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -k "not multicast" > gpurun_out/c9_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c9_pytest.log
tail -5 gpurun_out/c9_pytest.log
timeout 300 python profiles/profile_step.py --steps 1 --vae 0 2>&1 | grep "ms per" > gpurun_out/c9_ab.txt; cat gpurun_out/c9_ab.txt
timeout 900 python bench.py > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; cut -c1-300 gpurun_out/c9_bench.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c9_launches_warm.csv python profiles/profile_step.py --steps 1 > gpurun_out/c9_ncu1.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c9_smoke.log 2>&1; tail -4 gpurun_out/c9_smoke.log
ls -la gpurun_out | grep c9_
