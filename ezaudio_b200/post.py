"""Device-side waveform pre / post-processing around the path: the host-numpy statements of
api/ezaudio.py:147,198-203 and api/controlnet.py:119-136 as CUDA kernels behind the C-ABI, so a clip is uploaded once and
downloaded once.  float32 in, float32 out, identical arithmetic (true division by max|x| + 1e-9)."""
from __future__ import annotations

import torch

from . import _lib


def _dev(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise _lib.EzbError("waveform post-processing needs CUDA tensors (no CPU path)")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def prepare_wave(audio: torch.Tensor, out_samples: int | None = None, normalize: bool = True, gate: float = 0.0) -> torch.Tensor:
    """audio (B,T) fp32 cuda -> (B,out_samples): x / (max|x| + 1e-9) per clip, |x| <= gate zeroed, zero-padded / cropped."""
    a = audio.detach().to(torch.float32).contiguous()
    B, T = a.shape
    To = int(T if out_samples is None else out_samples)
    out = torch.empty(B, To, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().ezb_wave_prepare(_dev(a), _lib.ptr(a), _lib.ptr(out), B, T, To, int(bool(normalize)), float(gate or 0.0),
                                               torch.cuda.current_stream(a.device).cuda_stream))
    return out


def splice_wave(dst: torch.Tensor, src: torch.Tensor, start: int, n: int | None = None) -> torch.Tensor:
    """dst[start:start+n] = src[:n] in place (1-D fp32 cuda tensors); returns dst."""
    if dst.dim() != 1 or src.dim() != 1 or not dst.is_contiguous() or dst.dtype != torch.float32:
        raise ValueError("splice_wave: 1-D contiguous float32 tensors expected")
    s = src.detach().to(torch.float32).contiguous()
    n = int(s.numel() if n is None else n)
    if n > s.numel():
        raise ValueError("splice_wave: n exceeds the source length")
    with torch.cuda.device(dst.device):
        _lib.check(_lib.lib().ezb_wave_splice(_dev(dst), _lib.ptr(dst), dst.numel(), _lib.ptr(s), int(start), n,
                                              torch.cuda.current_stream(dst.device).cuda_stream))
    return dst


def to_pcm16(wav: torch.Tensor) -> torch.Tensor:
    """float waveform -> int16 PCM (round(x * 32768), saturated): the samples soundfile.write(..) stores in a default WAV."""
    w = wav.detach().to(torch.float32).contiguous()
    out = torch.empty(w.shape, dtype=torch.int16, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(_lib.lib().ezb_wave_to_pcm16(_dev(w), _lib.ptr(w), _lib.ptr(out), w.numel(), torch.cuda.current_stream(w.device).cuda_stream))
    return out
