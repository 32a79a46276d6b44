#!/bin/bash
# GPU call 5: attention with two MMA-issuing warps (bounded first), counters, A/B.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_attention_gpu.py -q -s > gpurun_out/c5_pytest_attn.log 2>&1; ARC=$?; echo "attention pytest exit $ARC" >> gpurun_out/c5_pytest_attn.log
tail -4 gpurun_out/c5_pytest_attn.log
if [ $ARC -eq 0 ]; then
timeout 300 python profiles/attn_bench.py 0 > gpurun_out/c5_attn_bench.txt 2>&1
timeout 300 python profiles/attn_bench.py 1 >> gpurun_out/c5_attn_bench.txt 2>&1; cat gpurun_out/c5_attn_bench.txt
timeout 300 python profiles/attn_dbg.py 1 > gpurun_out/c5_attn_dbg_mma2.txt 2>&1; cat gpurun_out/c5_attn_dbg_mma2.txt
timeout 600 python -m pytest tests/test_fold_gpu.py tests/test_dit_gpu.py tests/test_loop_gpu.py -q -s -k "attn_mma2 or dit_XL or loop" > gpurun_out/c5_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c5_pytest.log
tail -4 gpurun_out/c5_pytest.log
for o in "" "--opt attn_mma2=1" "--opt attn_mma2=1 --opt dhp80=1" "--opt attn_mma2=1 --opt attn5=1"; do
  timeout 300 python profiles/profile_step.py --steps 1 --vae 0 $o 2>&1 | grep "ms per" >> gpurun_out/c5_ab.txt
done
cat gpurun_out/c5_ab.txt
EZB_ATTN_MMA2=1 EZB_DHP80=1 timeout 900 python bench.py --no-cpu-baseline --no-extras > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; cut -c1-300 gpurun_out/c5_bench.json
fi
ls -la gpurun_out | grep c5_
