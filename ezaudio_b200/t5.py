"""T5 text encoder (the step before the denoiser path), backed by libezb200.so.

Drop-in for the `transformers.T5EncoderModel` object the reference builds at api/ezaudio.py:78-79 and calls at src/inference.py:38-50:
`text_encoder(input_ids=ids, attention_mask=mask).last_hidden_state`.  Same state-dict keys as `T5EncoderModel.state_dict()`.
Gated-GELU T5 v1.1 / flan-T5 configurations only (google/flan-t5-xl and -large are what the shipped configs name)."""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Dict

import torch

from . import _lib, weights
from .dit import PRECISIONS, _as_f32c


def relative_position_buckets(L: int, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """T5Attention._relative_position_bucket(key - query), bidirectional, with the same float32 torch ops as transformers (so the
    truncations at the bucket boundaries agree bit for bit): (L, L) int32, indexed [query, key]."""
    pos = torch.arange(L)
    rp = pos[None, :] - pos[:, None]
    nb = num_buckets // 2
    ret = (rp > 0).to(torch.long) * nb
    rp = torch.abs(rp)
    max_exact = nb // 2
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return (ret + torch.where(rp < max_exact, rp, large)).to(torch.int32).contiguous()


class T5EncoderModel:
    def __init__(self, config: dict, precision: str = "bf16", max_batch: int = 8, max_len: int = 100, device="cuda"):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {list(PRECISIONS)}")
        cfg = dict(config)
        if cfg.get("feed_forward_proj", "gated-gelu") != "gated-gelu":
            raise NotImplementedError("only gated-GELU T5 v1.1 / flan-T5 encoders (what the shipped EzAudio configs use)")
        self.config = cfg
        self.shapes = weights.t5_param_shapes(cfg)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.EzbError("ezaudio_b200 runs on CUDA devices only (no CPU path)")
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.max_batch, self.max_len = max_batch, max_len
        d = _lib.T5Desc(vocab_size=cfg["vocab_size"], d_model=cfg["d_model"], d_kv=cfg["d_kv"], num_heads=cfg["num_heads"], d_ff=cfg["d_ff"],
                        num_layers=cfg["num_layers"], num_buckets=cfg.get("relative_attention_num_buckets", 32),
                        max_distance=cfg.get("relative_attention_max_distance", 128), eps=cfg.get("layer_norm_epsilon", 1e-6),
                        max_batch=max_batch, max_len=max_len, precision=PRECISIONS[precision])
        self._buckets = {}
        self._graphs = {}     # (nb, L) -> (graph, static ids, static mask, static out): ~250 small launches replayed as one graph
        self.use_graphs = True
        self.h = C.c_void_p()
        with torch.cuda.device(self.dev_index):
            _lib.check(_lib.lib().ezb_t5_create(C.byref(self.h), C.byref(d), self.dev_index))

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                _lib.lib().ezb_t5_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    # transformers-style no-ops so that reference code paths (`.to(device)`, `.eval()`) keep working
    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        L = _lib.lib()
        with torch.cuda.device(self.dev_index):
            st = _lib.stream_ptr()
            for k, v in sd.items():
                if k not in self.shapes and k != "encoder.embed_tokens.weight":
                    if strict:
                        raise _lib.EzbError(f"unexpected state-dict key {k!r}")
                    continue
                t = _as_f32c(v).to(self.device)
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(L.ezb_t5_load_weight(self.h, k.encode(), _lib.ptr(t), shape, t.dim(), st))
                torch.cuda.current_stream().synchronize()  # `t` is a temporary
            _lib.check(L.ezb_t5_finalize_weights(self.h, st))
        return self

    @torch.no_grad()
    def __call__(self, input_ids=None, attention_mask=None, **unused):
        ids = input_ids.to(self.device).to(torch.int32).contiguous()
        B, L = ids.shape
        if attention_mask is None:
            attention_mask = torch.ones(B, L, dtype=torch.uint8)
        mask = attention_mask.to(self.device).to(torch.uint8).contiguous()
        if L not in self._buckets:
            self._buckets[L] = relative_position_buckets(L, self.config.get("relative_attention_num_buckets", 32),
                                                         self.config.get("relative_attention_max_distance", 128)).to(self.device)
        out = torch.empty(B, L, self.config["d_model"], device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.dev_index):
            for b0 in range(0, B, self.max_batch):
                nb = min(self.max_batch, B - b0)
                if self.use_graphs:
                    g, s_ids, s_mask, s_out = self._graph(nb, L)
                    s_ids.copy_(ids[b0:b0 + nb])
                    s_mask.copy_(mask[b0:b0 + nb])
                    g.replay()
                    out[b0:b0 + nb].copy_(s_out)
                else:
                    self._run(ids[b0:b0 + nb], mask[b0:b0 + nb], out[b0:b0 + nb], nb, L)
        return _Output(out)

    def _run(self, ids, mask, out, nb, L):
        _lib.check(_lib.lib().ezb_t5_forward(self.h, _lib.ptr(ids), _lib.ptr(mask), _lib.ptr(self._buckets[L]), C.c_void_p(out.data_ptr()), nb, L,
                                             _lib.stream_ptr()))

    def _graph(self, nb, L):
        key = (nb, L)
        if key not in self._graphs:
            s_ids = torch.zeros(nb, L, dtype=torch.int32, device=self.device)
            s_mask = torch.ones(nb, L, dtype=torch.uint8, device=self.device)
            s_out = torch.empty(nb, L, self.config["d_model"], device=self.device, dtype=torch.float32)
            self._run(s_ids, s_mask, s_out, nb, L)   # warm-up outside capture: tensor maps, function attributes
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run(s_ids, s_mask, s_out, nb, L)
            self._graphs[key] = (g, s_ids, s_mask, s_out)
        return self._graphs[key]

    forward = __call__


class _Output(SimpleNamespace):
    """`.last_hidden_state` and `[0]`, like transformers' BaseModelOutput."""

    def __init__(self, last_hidden_state):
        super().__init__(last_hidden_state=last_hidden_state)

    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]
