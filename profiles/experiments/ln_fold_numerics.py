"""CPU experiment (no GPU needed): does folding LayerNorm + AdaLN modulation into the consuming GEMM cost accuracy?

  standard : h = bf16(LN(x) * g + c) ;  out = h @ bf16(W)^T                      (what the library does today: a separate LN pass)
  folded   : A = bf16(x * g)         ;  out = rstd * (A @ bf16(W)^T - mu * u) + v,  u = g @ bf16(W)^T, v = c @ bf16(W)^T
             (the residual-stream epilogue writes A and per-row sums; the consumer's epilogue applies the per-row affine: no LN pass)

Every nn.Linear of the XL denoiser is emulated with bf16-rounded operands and fp32 accumulation; the rest stays fp32.  Prints the
error of the final DiT output against the unmodified reference's golden (tests/golden/dit_XL.npz) for both formulations."""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ezaudio_b200 import synth, weights  # noqa: E402
from oracle import ezaudio_oracle as O  # noqa: E402

MODE = "standard"
info = {}
bf = lambda t: t.to(torch.bfloat16).float()


def linear(inp, W, bias=None):
    Wb = bf(W)
    f = info.get(id(inp))
    if MODE == "folded" and f is not None:
        x, g, c = f
        mu = x.mean(-1, keepdim=True)
        rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-5)
        acc = bf(x * g) @ Wb.t()
        u, v = (g * torch.ones_like(x[..., :1, :])) @ Wb.t(), (c * torch.ones_like(x[..., :1, :])) @ Wb.t()
        out = rstd * (acc - mu * u) + v
    else:
        out = bf(inp) @ Wb.t()
    return out if bias is None else out + bias


def layer_norm(x, w, b, eps=1e-5):
    y = F.layer_norm(x, (x.shape[-1],), w, b, eps)
    if x.dim() == 3 and x.shape[-1] >= 1024:
        info[id(y)] = (x, w, b)
        keep.append(y)
    return y


def film_modulate(y, shift, scale):
    out = y * (1 + scale) + shift
    if id(y) in info:
        x, w, b = info[id(y)]
        info[id(out)] = (x, w * (1 + scale), b * (1 + scale) + shift)
        keep.append(out)
    return out


keep = []
shim = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F) if not k.startswith("_")})
shim.linear = linear
O.F = shim
O.layer_norm = layer_norm
O.film_modulate = film_modulate
O.linear = lambda x, sd, key: linear(x, sd[key + ".weight"], sd.get(key + ".bias"))

name = sys.argv[1] if len(sys.argv) > 1 else "dit_XL"
sys.path.insert(0, os.path.join(ROOT))
from tests import helpers  # noqa: E402
cfg, sd, inp, g = helpers.dit_case_inputs(name)
ref = torch.from_numpy(g["out"])
torch.set_num_threads(os.cpu_count())
for MODE in ("standard", "folded"):
    info.clear(); keep.clear()
    with torch.no_grad():
        out, _ = O.maskdit_forward(sd, cfg, inp["x"], inp["t"], inp["ctx"], inp["mask"], inp["gt"], inp["gt_mask"])
    e = (out - ref).abs()
    print(f"{name} [{MODE:8s}] max-abs {float(e.max()):.3e} mean-abs {float(e.mean()):.3e}")
