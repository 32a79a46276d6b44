#!/bin/bash
# GPU call 13: K/V-resident attention (bounded first).
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_attention_gpu.py -q -s -k "kv_resident" > gpurun_out/c13_pytest_attn.log 2>&1; ARC=$?; echo "attention pytest exit $ARC" >> gpurun_out/c13_pytest_attn.log
tail -4 gpurun_out/c13_pytest_attn.log
if [ $ARC -eq 0 ]; then
timeout 300 python profiles/attn_bench.py 0 > gpurun_out/c13_attn_bench.txt 2>&1
timeout 300 python profiles/attn_bench.py 2 >> gpurun_out/c13_attn_bench.txt 2>&1
timeout 300 python profiles/attn_bench.py 6 >> gpurun_out/c13_attn_bench.txt 2>&1; cat gpurun_out/c13_attn_bench.txt
timeout 600 python -m pytest tests/test_fold_gpu.py -q -s -k "attn_res" > gpurun_out/c13_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c13_pytest.log; tail -3 gpurun_out/c13_pytest.log
for o in "" "--opt attn_res=1" "" "--opt attn_res=1" "--opt attn_res=1 --opt attn_poly=1"; do
  timeout 300 python profiles/profile_step.py --steps 1 --vae 0 $o 2>&1 | grep "ms per" >> gpurun_out/c13_ab.txt
done
cat gpurun_out/c13_ab.txt
fi
ls -la gpurun_out | grep c13_
