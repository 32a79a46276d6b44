#!/bin/bash
# GPU call 10: EpiLinear spill fix (VAE convs, fallback linears), MLP-out on CTA-pair tiles.
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -k "not multicast" > gpurun_out/c10_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c10_pytest.log
tail -5 gpurun_out/c10_pytest.log
for o in "" "--opt mlp2_pair=1" "" "--opt mlp2_pair=1"; do
  timeout 300 python profiles/profile_step.py --steps 1 --vae 0 $o 2>&1 | grep -E "ms per|VAE" >> gpurun_out/c10_ab.txt
done
cat gpurun_out/c10_ab.txt
timeout 900 python bench.py > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; cut -c1-300 gpurun_out/c10_bench.json
ls -la gpurun_out | grep c10_
