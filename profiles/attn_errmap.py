"""Per-item (b*H + h, query tile) max-abs error map of one attention implementation against the fp32 reference: python profiles/attn_errmap.py <impl> [B H Lq Lk dh]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ezaudio_b200 import _lib  # noqa: E402

L = _lib.lib()
impl = int(sys.argv[1])
B, H, Lq, Lk, dh = (int(v) for v in sys.argv[2:7]) if len(sys.argv) > 6 else (8, 16, 500, 500, 72)
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randn(B, H, Lq, dh, device="cuda", generator=g) * 1.5
k = torch.randn(B, H, Lk, dh, device="cuda", generator=g) * 1.5
v = torch.randn(B, H, Lk, dh, device="cuda", generator=g)
ref = ((q @ k.transpose(-1, -2)) / math.sqrt(dh)).softmax(-1) @ v          # [B, H, Lq, dh]
dhp = 80 if (impl >= 100 and dh == 72) else (dh + 63) // 64 * 64
dvp, lkp = (dh + 15) // 16 * 16, (Lk + 7) // 8 * 8
q16 = torch.zeros(B * H, Lq, dhp, device="cuda", dtype=torch.bfloat16); q16[:, :, :dh] = q.reshape(B * H, Lq, dh).bfloat16()
k16 = torch.zeros(B * H, Lk, dhp, device="cuda", dtype=torch.bfloat16); k16[:, :, :dh] = k.reshape(B * H, Lk, dh).bfloat16()
vt = torch.zeros(B * H, dvp, lkp, device="cuda", dtype=torch.bfloat16); vt[:, :dh, :Lk] = v.reshape(B * H, Lk, dh).transpose(1, 2).bfloat16()
out = torch.full((B, Lq, H * dh), 77.0, device="cuda", dtype=torch.bfloat16)
_lib.check(L.ezb_test_attention(0, _lib.ptr(q16), _lib.ptr(k16), _lib.ptr(vt), None, _lib.ptr(out), B, H, Lq, Lk, dh, impl, _lib.stream_ptr()))
torch.cuda.synchronize()
o = out.float().reshape(B, Lq, H, dh).permute(0, 2, 1, 3)                # [B, H, Lq, dh]
err = (o - ref).abs().amax(-1)                                             # [B, H, Lq]
nqt = (Lq + 127) // 128
pad = nqt * 128 - Lq
e = torch.nn.functional.pad(err, (0, pad)).reshape(B * H, nqt, 128).amax(-1)   # [B*H, nqt]
items = e.reshape(-1)
bad = (items > 0.05).nonzero().flatten().tolist()
print(f"impl {impl}: {len(bad)} of {items.numel()} items wrong; first {bad[:24]}")
print("untouched (77.0) rows:", int((out.float() == 77.0).all(-1).sum()), "of", B * Lq)
for i in bad[:6]:
    bh, t = divmod(i, nqt)
    rows = err[bh // H, bh % H, t * 128:(t + 1) * 128]
    print(f" item {i} (cta {i % 148}, slot {i // 148}): max err {float(rows.max()):.3f}, bad rows {int((rows > 0.05).sum())} of {rows.numel()}, first bad row {int((rows > 0.05).nonzero()[0]) if (rows > 0.05).any() else -1}")
