"""VAE side of the hot path: `Autoencoder(embedding=z)` (src/modules/autoencoder_wrapper.py:74-77) ->
OobleckDecoder (src/modules/stable_vae/models/autoencoders.py:149-190), backed by libezb200.so."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib, weights
from .dit import PRECISIONS, _as_f32c


class OobleckDecoder:
    """OobleckDecoder (and, when `encoder_cfg` is given, OobleckEncoder + VAE bottleneck) on one ezb_vae handle."""

    def __init__(self, precision="bf16", max_batch=4, max_latent_len=512, device="cuda", encoder_cfg=None, **dec_cfg):
        self.shapes = weights.vae_decoder_param_shapes(dec_cfg)
        self.encoder_cfg = dict(encoder_cfg) if encoder_cfg else None
        if self.encoder_cfg is not None:
            weights.vae_encoder_param_shapes(self.encoder_cfg)  # validates the switches
            if (self.encoder_cfg["channels"], list(self.encoder_cfg["c_mults"]), list(self.encoder_cfg["strides"])) != \
                    (dec_cfg["channels"], list(dec_cfg["c_mults"]), list(dec_cfg["strides"])):
                raise NotImplementedError("encoder and decoder must mirror each other (ckpts/vae/config.json)")
        self.cfg = dict(dec_cfg)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.EzbError("ezaudio_b200 runs on CUDA devices only (no CPU path)")
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        n = len(dec_cfg["c_mults"])
        d = _lib.VaeDesc(latent_dim=dec_cfg["latent_dim"], channels=dec_cfg["channels"], out_channels=dec_cfg["out_channels"], n_stages=n,
                         max_batch=max_batch, max_latent_len=max_latent_len, precision=PRECISIONS[precision],
                         with_encoder=1 if encoder_cfg else 0, in_channels=encoder_cfg["in_channels"] if encoder_cfg else 0,
                         enc_latent_dim=encoder_cfg["latent_dim"] if encoder_cfg else 0)
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
                if not (k.startswith("decoder.") or (self.encoder_cfg is not None and k.startswith("encoder."))):
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

    def encode(self, audio: torch.Tensor, noise=None) -> torch.Tensor:
        """audio (B,1,T) -> latents (B,latent,T/hop): encoder + z = mean + (softplus(scale)+1e-4) * noise
        (`noise=None` draws torch.randn from the global RNG like the reference's vae_sample; pass False for the mean)."""
        if self.encoder_cfg is None:
            raise _lib.EzbError("this handle was created without encoder_cfg")
        a = _as_f32c(audio).to(self.device)
        B, ch, T = a.shape
        if ch != 1:
            raise ValueError("mono audio (B,1,T) expected")
        pad = (-T) % self.hop
        if pad:  # strided convs floor the length; zero-pad to a whole latent frame (the reference silently truncates)
            a = torch.nn.functional.pad(a, (0, pad))
            T += pad
        L = T // self.hop
        Cz = self.cfg["latent_dim"]
        if noise is None:
            noise = torch.randn(B, Cz, L, device=self.device, dtype=torch.float32)
        nz = None if noise is False else _as_f32c(noise).to(self.device)
        z = torch.empty(B, Cz, L, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.dev_index):
            for b0 in range(0, B, self.max_batch):
                nb = min(self.max_batch, B - b0)
                _lib.check(_lib.lib().ezb_vae_encode(self.h, _lib.ptr(a[b0:b0 + nb]), None if nz is None else C.c_void_p(nz[b0:b0 + nb].data_ptr()),
                                                     C.c_void_p(z[b0:b0 + nb].data_ptr()), nb, T, _lib.stream_ptr()))
        return z


class Autoencoder:
    """Call contract of src/modules/autoencoder_wrapper.py:7-83 for model_type 'stable_vae', quantization_first=True:
    exactly one of audio / embedding.  Decode is the hot path; encode = VAE encoder + bottleneck sampling."""

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
            return self.decoder.encode(audio)
        raise ValueError("Either audio or embedding must be provided.")

    forward = __call__
