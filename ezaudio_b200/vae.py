"""VAE side of the hot path: `Autoencoder(embedding=z)` (src/modules/autoencoder_wrapper.py:74-77) ->
OobleckDecoder (src/modules/stable_vae/models/autoencoders.py:149-190), backed by libezb200.so."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib, weights
from .dit import PRECISIONS, _as_f32c


class OobleckDecoder:
    def __init__(self, precision="bf16", max_batch=4, max_latent_len=512, device="cuda", **dec_cfg):
        self.shapes = weights.vae_decoder_param_shapes(dec_cfg)
        self.cfg = dict(dec_cfg)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.EzbError("ezaudio_b200 runs on CUDA devices only (no CPU path)")
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        n = len(dec_cfg["c_mults"])
        d = _lib.VaeDesc(latent_dim=dec_cfg["latent_dim"], channels=dec_cfg["channels"], out_channels=dec_cfg["out_channels"], n_stages=n,
                         max_batch=max_batch, max_latent_len=max_latent_len, precision=PRECISIONS[precision])
        for i in range(n):
            d.c_mults[i] = dec_cfg["c_mults"][i]
            d.strides[i] = dec_cfg["strides"][i]
        self.hop = 1
        for s in dec_cfg["strides"]:
            self.hop *= s
        self.max_batch = max_batch
        self.h = C.c_void_p()
        with torch.cuda.device(self.dev_index):
            _lib.check(_lib.lib().ezb_vae_create(C.byref(self.h), C.byref(d), self.dev_index))

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                _lib.lib().ezb_vae_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        """Accepts the reference VAE state-dict after its 'autoencoder.' prefix strip (stable_vae/__init__.py:25-31);
        only `decoder.*` entries are consumed."""
        L = _lib.lib()
        with torch.cuda.device(self.dev_index):
            st = _lib.stream_ptr()
            for k, v in sd.items():
                if not k.startswith("decoder."):
                    continue
                t = _as_f32c(v).to(self.device)
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(L.ezb_vae_load_weight(self.h, k.encode(), _lib.ptr(t), shape, t.dim(), st))
            torch.cuda.current_stream().synchronize()
            _lib.check(L.ezb_vae_finalize_weights(self.h, st))
        return self

    def __call__(self, z: torch.Tensor) -> torch.Tensor:
        z = _as_f32c(z).to(self.device)
        B, Cz, L = z.shape
        wav = torch.empty(B, 1, L * self.hop, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.dev_index):
            for b0 in range(0, B, self.max_batch):
                nb = min(self.max_batch, B - b0)
                _lib.check(_lib.lib().ezb_vae_decode(self.h, _lib.ptr(z[b0:b0 + nb]), C.c_void_p(wav[b0:b0 + nb].data_ptr()), nb, L, _lib.stream_ptr()))
        return wav

    forward = __call__


class Autoencoder:
    """Call contract of src/modules/autoencoder_wrapper.py:7-83 for model_type 'stable_vae', quantization_first=True:
    exactly one of audio / embedding.  Decode is the hot path; encode (VAE encoder + bottleneck sampling, SURVEY 8f
    row 1) is not built yet and raises."""

    def __init__(self, decoder: OobleckDecoder):
        self.decoder = decoder

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def __call__(self, audio=None, embedding=None):
        if embedding is not None:
            return self.decoder(embedding)
        if audio is not None:
            raise NotImplementedError("VAE encode path (OobleckEncoder + VAEBottleneck) is the next SURVEY 8(f) row; not built yet")
        raise ValueError("Either audio or embedding must be provided.")

    forward = __call__
