#!/bin/bash
# GPU call 16: source-level ncu capture of the attention kernel (self and cross, XL shapes)
set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn4 -s 3 -c 1 -f -o gpurun_out/c16_attn_self python profiles/attn_one.py self 3 > gpurun_out/c16_ncu_self.log 2>&1; tail -3 gpurun_out/c16_ncu_self.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn4 -s 3 -c 1 -f -o gpurun_out/c16_attn_cross python profiles/attn_one.py cross 3 > gpurun_out/c16_ncu_cross.log 2>&1; tail -3 gpurun_out/c16_ncu_cross.log
ls -la gpurun_out | grep c16_
