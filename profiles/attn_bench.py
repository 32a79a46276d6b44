"""Attention micro-benchmark through the C-ABI test hook + CTA-0 cycle counters."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ezaudio_b200 import _lib  # noqa: E402

L = _lib.lib()
L.ezb_set_option(b"gemm_debug", 1)
V4 = 1
if len(sys.argv) > 1:
    V4 = int(sys.argv[1])
    L.ezb_set_option(b"attn4", V4)
if len(sys.argv) > 2:
    L.ezb_set_option(b"attn_poly", int(sys.argv[2]))
    print("attn_poly =", sys.argv[2])
    print("attn4 =", sys.argv[1])


def run(B, H, Lq, Lk, dh, masked, label, reps=20):
    dhp, dvp, lkp = (dh + 63) // 64 * 64, (dh + 15) // 16 * 16, (Lk + 7) // 8 * 8
    q = torch.randn(B * H, Lq, dhp, device="cuda").bfloat16()
    k = torch.randn(B * H, Lk, dhp, device="cuda").bfloat16()
    vt = torch.randn(B * H, dvp, lkp, device="cuda").bfloat16()
    q[:, :, dh:] = 0
    k[:, :, dh:] = 0
    mask = None
    if masked:
        mask = torch.zeros(B, Lk, dtype=torch.uint8, device="cuda")
        mask[:, :20] = 1
    out = torch.empty(B, Lq, H * dh, device="cuda", dtype=torch.bfloat16)
    args = (0, _lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), _lib.ptr(mask), _lib.ptr(out), B, H, Lq, Lk, dh, 1, _lib.stream_ptr())
    for _ in range(3):
        _lib.check(L.ezb_test_attention(*args))
    dbg = (C.c_ulonglong * 8)()
    L.ezb_debug_read(dbg)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        L.ezb_test_attention(*args)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    L.ezb_debug_read(dbg)
    d = [v / reps for v in dbg]
    fl = 4.0 * B * H * Lq * Lk * dh
    if V4:
        print(f"{label:18s} B{B} H{H} Lq{Lq} Lk{Lk} dh{dh}: {ms * 1e3:7.1f} us {fl / ms / 1e9:6.1f} TF/s | softmax g0: wait_S {d[0]:.0f} wait_O {d[1]:.0f} max+exp {d[2]:.0f} "
              f"st+arrive {d[3]:.0f} out {d[4]:.0f} total {d[5]:.0f} | mma: wait_P {d[6]:.0f} wait_QKV {d[7]:.0f}")
    else:
        print(f"{label:18s} B{B} H{H} Lq{Lq} Lk{Lk} dh{dh}: {ms * 1e3:7.1f} us {fl / ms / 1e9:6.1f} TF/s | softmax: wait_S {d[0]:.0f} wait_O {d[1]:.0f} "
              f"barrier {d[2]:.0f} total {d[3]:.0f} | mma: wait_kv {d[4]:.0f} wait_P {d[5]:.0f} total {d[6]:.0f}")


run(8, 16, 500, 500, 72, False, "self XL")
run(8, 16, 500, 100, 72, True, "cross XL")
run(4, 16, 1500, 1500, 72, False, "self XL 30s")
run(8, 16, 256, 256, 64, False, "self L")
