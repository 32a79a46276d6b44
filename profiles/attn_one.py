"""One attention configuration through the C-ABI test hook (for ncu source-level captures):
   python profiles/attn_one.py self|cross [reps] [opt=value ...]   -> XL shapes, B = 8, H = 16, L = 500 (cross: 100 masked keys), dhp = 80."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ezaudio_b200 import _lib  # noqa: E402

L = _lib.lib()
kind = sys.argv[1] if len(sys.argv) > 1 else "self"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    _lib.check(L.ezb_set_option(k.encode(), int(v)))
B, H, Lq, dh, dhp, dvp = 8, 16, 500, 72, 80, 80
Lk = 500 if kind == "self" else 100
lkp = (Lk + 7) // 8 * 8
q = torch.randn(B * H, Lq, dhp, device="cuda").bfloat16()
k = torch.randn(B * H, Lk, dhp, device="cuda").bfloat16()
vt = torch.randn(B * H, dvp, lkp, device="cuda").bfloat16()
q[:, :, dh:] = 0
k[:, :, dh:] = 0
mask = None
if kind != "self":
    mask = torch.zeros(B, Lk, dtype=torch.uint8, device="cuda")
    mask[:, :20] = 1
out = torch.empty(B, Lq, H * dh, device="cuda", dtype=torch.bfloat16)
args = (0, _lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), _lib.ptr(mask), _lib.ptr(out), B, H, Lq, Lk, dh, 101, _lib.stream_ptr())
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    _lib.check(L.ezb_test_attention(*args))
t0.record()
for _ in range(reps):
    L.ezb_test_attention(*args)
t1.record()
torch.cuda.synchronize()
print(f"{kind}: {t0.elapsed_time(t1) / reps * 1e3:.1f} us per launch")
