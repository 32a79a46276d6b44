"""Public API -- same classes, method signatures, defaults and return types as the reference
(api/ezaudio.py:31-207 `EzAudio`, api/controlnet.py:31-161 `EzAudio_ControlNet`), hosted on the CUDA library.

Extensions (all optional, keyword-only): `text` may be a list of prompts (batched; returns a list of waveforms);
`text_encoder=` injects a callable `(list[str]) -> (emb (B,Lc,ctx) , mask (B,Lc))` standing in for flan-T5 (the image has no
network, so T5 weights cannot be fetched -- BASELINE configs use cached embeddings); `ckpt_path="synthetic:<seed>"` builds
the deterministic random checkpoint of `weights.synthetic_state_dict` instead of reading a file.
"""
from __future__ import annotations

import os
import random
from typing import Callable, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib, config, post, weights
from ._lib import EzbError
from .dit import DiTControlNet, MaskDiT
from .inference import inference
from .scheduler import DDIMScheduler
from .vae import Autoencoder, OobleckDecoder

MAX_SEED = np.iinfo(np.int32).max


class SyntheticTextEncoder:
    """Cached-T5 stand-in: deterministic N(0,1) embeddings keyed by the prompt string; "" -> only the first
    (EOS) token is valid, like the tokenizer's output for the empty negative prompt."""

    def __init__(self, ctx_dim: int, max_length: int = 100):
        self.ctx_dim, self.max_length = ctx_dim, max_length

    def __call__(self, prompts: Sequence[str]):
        embs, masks = [], []
        for p in prompts:
            seed = int.from_bytes(p.encode()[:8].ljust(8, b"\0"), "little") % (2 ** 31) + 7 * len(p)
            g = torch.Generator().manual_seed(seed)
            embs.append(torch.randn(1, self.max_length, self.ctx_dim, generator=g))
            n = 1 if p == "" else min(self.max_length, 2 + len(p.split()) + len(p) // 6)
            m = torch.zeros(1, self.max_length, dtype=torch.bool)
            m[0, :n] = True
            masks.append(m)
        return torch.cat(embs, 0), torch.cat(masks, 0)


def _load_audio(path: str, sr: int) -> np.ndarray:
    """librosa.load(path, sr=sr) stand-in (mono float32, resampled), api/ezaudio.py:146.  librosa / soundfile / torchcodec are
    not in this image: WAV is read with scipy, resampling is polyphase (scipy.signal.resample_poly)."""
    from math import gcd

    from scipy.io import wavfile
    from scipy.signal import resample_poly
    fs, data = wavfile.read(path)
    if data.dtype.kind == "i":
        data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
    elif data.dtype.kind == "u":
        data = (data.astype(np.float32) - 128.0) / 128.0
    data = data.astype(np.float32)
    if data.ndim == 2:
        data = data.mean(axis=1)
    if fs != sr:
        g = gcd(int(fs), int(sr))
        data = resample_poly(data, sr // g, fs // g).astype(np.float32)
    return data


class HashTokenizer:
    """Offline stand-in for T5Tokenizer (no sentencepiece vocabulary on disk, no network): words -> stable ids in [2, vocab), EOS = 1,
    pad = 0; same call signature / outputs as the tokenizer call at src/inference.py:39-41."""

    def __init__(self, vocab_size: int = 32128):
        self.vocab_size = vocab_size

    def __call__(self, text, max_length=100, padding="max_length", truncation=True, return_tensors="pt"):
        from types import SimpleNamespace
        import zlib
        text = [text] if isinstance(text, str) else list(text)
        ids = torch.zeros(len(text), max_length, dtype=torch.long)
        mask = torch.zeros(len(text), max_length, dtype=torch.long)
        for i, t in enumerate(text):
            toks = [2 + zlib.crc32(w.encode()) % (self.vocab_size - 2) for w in t.lower().split()][: max_length - 1] + [1]
            ids[i, : len(toks)] = torch.tensor(toks)
            mask[i, : len(toks)] = 1
        return SimpleNamespace(input_ids=ids, attention_mask=mask)


class NativeTextEncoder:
    """prompts -> (embeddings, mask) with a tokenizer and the native T5 encoder (ezaudio_b200.t5), i.e. src/inference.py:38-50."""

    def __init__(self, tokenizer, encoder, max_length: int = 100, device="cuda"):
        self.tokenizer, self.encoder, self.max_length, self.device = tokenizer, encoder, max_length, device

    def __call__(self, prompts: Sequence[str]):
        tb = self.tokenizer(list(prompts), max_length=self.max_length, padding="max_length", truncation=True, return_tensors="pt")
        ids, mask = tb.input_ids.to(self.device), tb.attention_mask.to(self.device).bool()
        return self.encoder(input_ids=ids, attention_mask=mask).last_hidden_state, mask


def save_wav(path: str, audio, sr: int = 24000) -> None:
    """soundfile.write(path, audio, sr) stand-in for the step after the path (t2a_demo.py:13,20, controlnet_demo.py:15): float32 mono WAV
    written with scipy (soundfile is not in this image).  Accepts the (sr, ndarray) tuple the API methods return as `audio`."""
    from scipy.io import wavfile
    if isinstance(audio, tuple):
        sr, audio = audio
    a = np.asarray(audio, dtype=np.float32).reshape(-1)
    if not np.isfinite(a).all():
        raise ValueError("save_wav: non-finite samples")
    wavfile.write(path, int(sr), a)


def _load_t5(name: str, device, precision: str = "bf16", max_length: int = 100):
    """api/ezaudio.py:78-79.  transformers is used for the tokenizer and for reading the checkpoint (CPU, I/O only); the encoder that runs
    is the native one.  Returns (None, None) when the checkpoint is not on disk (there is no network here)."""
    try:
        from transformers import T5EncoderModel as HFT5
        from transformers import T5Tokenizer
        tok = T5Tokenizer.from_pretrained(name, local_files_only=True)
        hf = HFT5.from_pretrained(name, local_files_only=True)
    except Exception:
        return None, None
    from .t5 import T5EncoderModel
    c = hf.config
    cfg = dict(vocab_size=c.vocab_size, d_model=c.d_model, d_kv=c.d_kv, num_heads=c.num_heads, d_ff=c.d_ff, num_layers=c.num_layers,
               relative_attention_num_buckets=c.relative_attention_num_buckets,
               relative_attention_max_distance=getattr(c, "relative_attention_max_distance", 128), layer_norm_epsilon=c.layer_norm_epsilon,
               feed_forward_proj=c.feed_forward_proj)
    enc = T5EncoderModel(cfg, precision=precision, max_batch=8, max_len=max_length, device=device).load_state_dict(hf.state_dict())
    return tok, enc


def _state_dict(path, shapes, key):
    if isinstance(path, str) and path.startswith("synthetic"):
        seed = int(path.split(":")[1]) if ":" in path else 0
        return weights.synthetic_state_dict(shapes, seed)
    if path is None or not os.path.exists(path):
        raise FileNotFoundError(f"checkpoint {path!r} not found (no network here: pass a local file or 'synthetic:<seed>')")
    sd = torch.load(path, map_location="cpu")
    return sd[key] if key in sd else sd


class _Base:
    def _text_embeds(self, prompts: List[str], neg: List[str]):
        if self.encode_text is None:
            raise RuntimeError("no text encoder available (flan-T5 weights are not on disk and there is no network); pass "
                               "text_encoder=<callable> to the constructor, e.g. ezaudio_b200.api.SyntheticTextEncoder")
        e, m = self.encode_text(prompts)
        ue, um = self.encode_text(neg)
        return e, m, ue, um

    def _make_text_encoder(self, text_encoder, params, device):
        self.tokenizer = self.text_encoder = None
        if text_encoder is not None:
            return text_encoder
        ml = params["text_encoder"]["max_length"]
        tok, enc = _load_t5(params["text_encoder"]["model"], device, max_length=ml)
        if tok is None:
            return None
        self.tokenizer, self.text_encoder = tok, enc
        return NativeTextEncoder(tok, enc, ml, device)


class EzAudio(_Base):
    """api/ezaudio.py:31."""

    def __init__(self, model_name, ckpt_path=None, vae_path=None, device="cuda", *, text_encoder: Optional[Callable] = None,
                 precision: str = "bf16", max_batch: int = 4, max_length_s: float = 10.0, config_path=None, vae_config_path=None):
        self.device = device
        self.params = config.load_params(model_name, config_path)
        p = self.params
        latent_sr = p["autoencoder"]["latent_sr"]
        max_len = int(round(max_length_s * latent_sr))
        self.noise_scheduler = DDIMScheduler(**p["diff"])
        self.unet = MaskDiT(precision=precision, max_batch=2 * max_batch, max_len=max_len, max_ctx_len=p["text_encoder"]["max_length"],
                            max_timesteps=1000, device=device, **p["model"])
        self.unet.load_state_dict(_state_dict(ckpt_path, weights.dit_param_shapes(p["model"]), "model"))
        dcfg, ecfg = config.load_vae_decoder_config(vae_config_path), config.load_vae_encoder_config(vae_config_path)
        dec = OobleckDecoder(precision=precision, max_batch=max_batch, max_latent_len=max_len, device=device, encoder_cfg=ecfg, **dcfg)
        vshapes = dict(weights.vae_decoder_param_shapes(dcfg))
        vshapes.update(weights.vae_encoder_param_shapes(ecfg))
        vsd = _state_dict(vae_path, vshapes, "state_dict")
        vsd = {(k[len("autoencoder."):] if k.startswith("autoencoder.") else k): v for k, v in vsd.items()}  # stable_vae/__init__.py:25-31
        dec.load_state_dict(vsd)
        self.autoencoder = Autoencoder(dec)
        self.encode_text = self._make_text_encoder(text_encoder, p, device)

    def generate_audio(self, text, length=10, guidance_scale=5, guidance_rescale=0.75, ddim_steps=100, eta=1, random_seed=None,
                       randomize_seed=False):
        """api/ezaudio.py:101-130.  Returns (sr, float32 waveform); a list of prompts returns (sr, [waveforms])."""
        batched = not isinstance(text, str)
        prompts = list(text) if batched else [text]
        length = length * self.params["autoencoder"]["latent_sr"]
        if all(t == "" for t in prompts):
            guidance_scale = None
            print("empyt input")
        if randomize_seed:
            random_seed = random.randint(0, MAX_SEED)
        embeds = self._text_embeds(prompts, [""])
        pred = inference(self.autoencoder, self.unet, None, None, None, None, self.params, self.noise_scheduler, prompts, None,
                         int(length), guidance_scale, guidance_rescale, ddim_steps, eta, random_seed, self.device, text_embeds=embeds)
        pred = pred.cpu().numpy()
        sr = self.params["autoencoder"]["sr"]
        if batched:
            return sr, [pred[i, 0] for i in range(pred.shape[0])]
        return sr, pred.squeeze(0).squeeze(0)

    def editing_audio(self, text, boundary, gt_file, mask_start, mask_length, guidance_scale=3.5, guidance_rescale=0, ddim_steps=100,
                      eta=1, random_seed=None, randomize_seed=False):
        """api/ezaudio.py:132-207 (crop -> VAE encode -> masked sampling -> paste -> decode -> splice)."""
        sr = self.params["autoencoder"]["sr"]
        if text == "":
            guidance_scale = None
            print("empyt input")
        mask_end = mask_start + mask_length
        gt_raw = _load_audio(gt_file, sr) if isinstance(gt_file, str) else np.asarray(gt_file, dtype=np.float32)
        audio_length = len(gt_raw) / sr
        mask_start = min(mask_start, audio_length)
        n_total = len(gt_raw)
        if mask_end > audio_length:  # outpainting: zero padding up to the end of the mask
            n_total += round((mask_end - audio_length) * sr)
            audio_length = n_total / sr
        # the clip goes to the device ONCE: peak-normalise + pad there (ezb_wave_prepare = api/ezaudio.py:147,152-154 on the device)
        output_audio = post.prepare_wave(torch.from_numpy(gt_raw).to(self.device).unsqueeze(0), n_total, normalize=True)[0]
        boundary = min((mask_end - mask_start) / 2, boundary)
        start_idx = max(mask_start - boundary, 0)
        end_idx = min(mask_end + boundary, audio_length)
        mask_start -= start_idx
        mask_end -= start_idx
        s0, s1 = round(start_idx * sr), round(end_idx * sr)
        gt_t = output_audio[s0:s1].clone().view(1, 1, -1)
        gt_latent = self.autoencoder(audio=gt_t)  # OobleckEncoder + stochastic VAE bottleneck (global RNG, bottleneck.py:69)
        B, D, L = gt_latent.shape
        gt_mask = torch.zeros(B, D, L, device=self.device)
        latent_sr = self.params["autoencoder"]["latent_sr"]
        gt_mask[:, :, round(mask_start * latent_sr):round(mask_end * latent_sr)] = 1
        gt_mask = gt_mask.bool()
        if randomize_seed:
            random_seed = random.randint(0, MAX_SEED)
        embeds = self._text_embeds([text], [""])
        pred = inference(self.autoencoder, self.unet, gt_latent, gt_mask, None, None, self.params, self.noise_scheduler, [text], None, L,
                         guidance_scale, guidance_rescale, ddim_steps, eta, random_seed, self.device, text_embeds=embeds)
        # trim + paste on the device (ezb_wave_splice = api/ezaudio.py:198-203), one download of the finished clip
        n = min(round((end_idx - start_idx) * sr), pred.shape[-1], n_total - s0)
        post.splice_wave(output_audio, pred[0, 0], s0, n)
        return sr, output_audio.cpu().numpy()


def energy_condition(audio: torch.Tensor, hop_size=240, window_size=1920, padding="reflect", min_db=-60, norm=True, quantize_levels=None,
                     **unused):
    """EnergyExtractor + Conditioner (src/models/conditions/energy.py:19-56, condition_wrapper.py:26-42): (B,T) -> (B,1,T/hop).
    One CUDA kernel (ezb_energy_condition); no torch fallback."""
    if padding != "reflect":
        raise NotImplementedError("energy conditioner: only padding='reflect' (the shipped config)")
    if not audio.is_cuda:
        raise EzbError("energy_condition needs a CUDA tensor")
    a = audio.detach().to(torch.float32).contiguous()
    B, T = a.shape
    out = torch.empty(B, 1, T // hop_size, dtype=torch.float32, device=a.device)
    _lib.check(_lib.lib().ezb_energy_condition(a.device.index or 0, a.data_ptr(), out.data_ptr(), B, T, int(hop_size), int(window_size), float(min_db),
                                               int(bool(norm)), int(quantize_levels or 0), torch.cuda.current_stream(a.device).cuda_stream))
    return out


class EzAudio_ControlNet(_Base):
    """api/controlnet.py:31."""

    def __init__(self, model_name, ckpt_path=None, controlnet_path=None, vae_path=None, device="cuda", *,
                 text_encoder: Optional[Callable] = None, precision: str = "bf16", max_batch: int = 4, config_path=None,
                 vae_config_path=None, params: Optional[dict] = None):
        self.device = device
        self.params = params if params is not None else config.load_params(model_name, config_path, config.BUILTIN_CONTROLNET)
        p = self.params
        max_len = 10 * p["autoencoder"]["latent_sr"]  # the reference ControlNet API is hard-wired to 10 s (api/controlnet.py:131-138)
        self.noise_scheduler = DDIMScheduler(**p["diff"])
        kw = dict(precision=precision, max_batch=2 * max_batch, max_len=max_len, max_ctx_len=p["text_encoder"]["max_length"], max_timesteps=1000,
                  device=device)
        self.unet = MaskDiT(**kw, **p["model"])
        sd = _state_dict(ckpt_path, weights.dit_param_shapes(p["model"]), "model")
        self.unet.load_state_dict(sd)
        self.controlnet = DiTControlNet(**kw, **p["model"], **p["controlnet"])
        csd = _state_dict(controlnet_path, weights.controlnet_param_shapes(p["model"], p["controlnet"]), "model")
        self.controlnet.load_state_dict(csd, mask_embed=sd["mask_embed"])
        dcfg = config.load_vae_decoder_config(vae_config_path)
        dec = OobleckDecoder(precision=precision, max_batch=max_batch, max_latent_len=max_len, device=device, **dcfg)
        vsd = _state_dict(vae_path, weights.vae_decoder_param_shapes(dcfg), "state_dict")
        vsd = {(k[len("autoencoder."):] if k.startswith("autoencoder.") else k): v for k, v in vsd.items()}
        dec.load_state_dict(vsd)
        self.autoencoder = Autoencoder(dec)
        if p["conditioner"]["condition_type"] != "energy":
            raise NotImplementedError("only the shipped energy conditioner")
        self.encode_text = self._make_text_encoder(text_encoder, p, device)

    def generate_audio(self, text, audio_path, surpass_noise=0, guidance_scale=3.5, guidance_rescale=0, ddim_steps=50, eta=1,
                       conditioning_scale=1, random_seed=None, randomize_seed=False):
        """api/controlnet.py:113-161.  `audio_path` may also be a float32 numpy waveform at the model sample rate."""
        sr = self.params["autoencoder"]["sr"]
        gt = _load_audio(audio_path, sr) if isinstance(audio_path, str) else np.asarray(audio_path, dtype=np.float32)
        original_length = len(gt)
        num_samples = int(10 * sr)
        audio_frames = round(num_samples / sr * self.params["autoencoder"]["latent_sr"])
        # normalise, noise-gate and pad / crop to 10 s on the device (ezb_wave_prepare = api/controlnet.py:119-136)
        gt_audio = post.prepare_wave(torch.from_numpy(gt).to(self.device).unsqueeze(0), num_samples, normalize=True, gate=float(surpass_noise or 0))
        # the reference encodes gt_audio only to read its latent SHAPE (api/controlnet.py:141-142): (1, 128, audio_frames)
        cond_kw = {k: v for k, v in self.params["conditioner"].items() if k != "condition_type"}
        condition = energy_condition(gt_audio, **cond_kw)
        if randomize_seed:
            random_seed = random.randint(0, MAX_SEED)
        batched = not isinstance(text, str)
        prompts = list(text) if batched else [text]
        condition = condition.expand(len(prompts), -1, -1)
        embeds = self._text_embeds(prompts, [""])
        pred = inference(self.autoencoder, self.unet, None, None, None, None, self.params, self.noise_scheduler, prompts, None, audio_frames,
                         guidance_scale, guidance_rescale, ddim_steps, eta, random_seed, self.device, text_embeds=embeds,
                         controlnet=self.controlnet, condition=condition, conditioning_scale=conditioning_scale)
        pred = pred.cpu().numpy()
        if batched:
            return sr, [pred[i, 0][:original_length] for i in range(pred.shape[0])]
        return sr, pred.squeeze(0).squeeze(0)[:original_length]
