#!/bin/bash
# GPU call 19: attention generation 6 v2 (Q prefetch, TMA-store write-out at item end, 3-deep rings) + wave_out with batched loads
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -q -x -k "gen6" > gpurun_out/c19_pytest_attn.log 2>&1; RC=$?; echo "pytest exit $RC" >> gpurun_out/c19_pytest_attn.log
tail -15 gpurun_out/c19_pytest_attn.log
for m in 0 1 3 5 7; do timeout 200 python profiles/attn_bench.py 0 $m 2>&1 | grep -v "impl   5" >> gpurun_out/c19_attn_bench.txt; done; cat gpurun_out/c19_attn_bench.txt
timeout 600 python -m pytest tests/test_vae_gpu.py -q > gpurun_out/c19_pytest_vae.log 2>&1; echo "vae pytest exit $?" >> gpurun_out/c19_pytest_vae.log; tail -3 gpurun_out/c19_pytest_vae.log
if [ $RC -eq 0 ]; then
timeout 600 python profiles/ab_sweep.py "" "attn6=1" "attn6=3" "attn6=5" "attn6=7" "" "attn6=1" > gpurun_out/c19_ab.txt 2> gpurun_out/c19_ab.err; cat gpurun_out/c19_ab.txt; tail -3 gpurun_out/c19_ab.err
fi
timeout 300 python profiles/profile_step.py --steps 1 2>&1 | grep -E "VAE|ms per" > gpurun_out/c19_step.txt; cat gpurun_out/c19_step.txt
ls -la gpurun_out | grep c19_
