"""Where does the attention kernel's time go?  Times attn4 (self, XL shapes) with parts of its work removed (option attn_dbg; results are
garbage, only the durations mean something): 1 no exp2 (MUFU), 2 no S load from tensor memory, 4 no P store, 8 no P V MMAs, 16 no S MMAs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ezaudio_b200 import _lib  # noqa: E402

import ctypes as C
L = _lib.lib()
_lib.check(L.ezb_set_option(b"gemm_debug", 1))
_lib.check(L.ezb_set_option(b"attn6", 0))   # the profiling instantiation belongs to the generation-4 kernel
B, H, Lq, Lk, dh = 8, 16, 500, 500, 72
dhp, dvp, lkp = 128, 80, 504
q = torch.randn(B * H, Lq, dhp, device="cuda").bfloat16()
k = torch.randn(B * H, Lk, dhp, device="cuda").bfloat16()
vt = torch.randn(B * H, dvp, lkp, device="cuda").bfloat16()
q[:, :, dh:] = 0
k[:, :, dh:] = 0
out = torch.empty(B, Lq, H * dh, device="cuda", dtype=torch.bfloat16)
args = (0, _lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), None, _lib.ptr(out), B, H, Lq, Lk, dh, 1, _lib.stream_ptr())
for mode, label in ((0, "full kernel"), (1, "no exp2"), (2, "no S load"), (4, "no P store"), (3, "no exp2, no S load"), (7, "no exp2 / S load / P store"),
                    (8, "no PV MMA"), (16, "no S MMA"), (24, "no MMA at all"), (31, "nothing but the hand-offs")):
    _lib.check(L.ezb_set_option(b"attn_dbg", mode | 32))   # 32: CTA-0 cycle counters; the profiling instantiation (DBG = 1) also runs for mode 0
    for _ in range(3):
        _lib.check(L.ezb_test_attention(*args))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(20):
        L.ezb_test_attention(*args)
    t1.record()
    torch.cuda.synchronize()
    d = (C.c_ulonglong * 8)()
    L.ezb_debug_read(d)
    print(f"attn_dbg {mode:2d} ({label:28s}): {t0.elapsed_time(t1) / 20 * 1e3:6.1f} us | CTA 0 cycles: softmax g0 wait_S {d[0]} write-out {d[1]} loop {d[2]} | "
          f"MMA wait_P {d[3]} wait_V {d[4]} wait_QK {d[5]} total {d[7]} | TMA wait_empty {d[6]}")
_lib.check(L.ezb_set_option(b"attn_dbg", 0))
