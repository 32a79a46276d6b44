#!/bin/bash
# QKV GEMM with parts of its epilogue removed (option heads_dbg; garbage results): which part is on the critical path?
for m in 0 16 1 2 4 8 3 12 15; do
  python profiles/profile_step.py --steps 1 --vae 0 --opt heads_dbg=$m 2>&1 | grep "ms per"
done
