"""Synthetic inputs of SURVEY.md section 8(d): cached-T5 stand-ins, latents, masks, tiny configs.

Shared by tests, bench.py and the golden generator so every party sees the same bits
(CPU generators, fixed seeds)."""
from __future__ import annotations

import copy
from typing import Dict

import torch

XL_MODEL = dict(mae=True, mae_prob=0.25, mask_ratio=[0.25, 1.0], mask_span=10, img_size=500, patch_size=1,
                in_chans=257, out_chans=128, input_type="1d", embed_dim=1152, depth=28, num_heads=16,
                mlp_ratio=4.0, qkv_bias=False, qk_scale=None, qk_norm="layernorm", norm_layer="layernorm",
                act_layer="geglu", context_norm=True, use_checkpoint=True, time_fusion="ada_sola_bias",
                ada_sola_rank=36, ada_sola_alpha=36, cls_dim=None, context_dim=2048, context_fusion="cross",
                context_max_length=None, context_pe_method="none", pe_method="none", rope_mode="shared",
                use_conv=True, skip=True, skip_norm=True)  # ckpts/ezaudio-xl.yml:5-37
L_MODEL = dict(XL_MODEL, embed_dim=1024, depth=24, ada_sola_rank=32, ada_sola_alpha=32,
               context_dim=1024)  # ckpts/ezaudio-l.yml
CONTROLNET = dict(cond_in=1, cond_blocks=[64, 128], cond_mask=True, cond_mask_prob=0.25,
                  cond_mask_ratio=[0.25, 0.50], cond_mask_span=10)  # ckpts/controlnet/energy_l.yml:38-44
VAE_DECODER = dict(out_channels=1, channels=128, c_mults=[1, 2, 4, 8], strides=[2, 4, 6, 10],
                   latent_dim=128, use_snake=True, final_tanh=False)  # ckpts/vae/config.json:18-28


VAE_ENCODER = dict(in_channels=1, channels=128, c_mults=[1, 2, 4, 8], strides=[2, 4, 6, 10], latent_dim=256,
                   use_snake=True)  # ckpts/vae/config.json:7-16


def tiny_model(head_dim: int = 72, heads: int = 2, depth: int = 4, ctx_dim: int = 64, rank: int = 4) -> Dict:
    """Same architecture switches as the shipped configs, small dims (head_dim 72 like XL or 64 like L)."""
    return dict(XL_MODEL, embed_dim=head_dim * heads, num_heads=heads, depth=depth, context_dim=ctx_dim,
                ada_sola_rank=rank, ada_sola_alpha=rank, img_size=64)


def tiny_vae(channels: int = 16) -> Dict:
    return dict(VAE_DECODER, channels=channels)


def tiny_vae_encoder(channels: int = 16) -> Dict:
    return dict(VAE_ENCODER, channels=channels)


def model_cfg(name: str) -> Dict:
    return copy.deepcopy({"xl": XL_MODEL, "l": L_MODEL}[name])


def synth_latents(B: int, L: int, C: int = 128, seed: int = 2024) -> torch.Tensor:
    """x0 ~ N(0,1), one generator per prompt (seed + i) so batch-B equals B reference runs."""
    return torch.cat([torch.randn(1, C, L, generator=torch.Generator().manual_seed(seed + i)) for i in range(B)], 0)


def synth_context(B: int, Lc: int, ctx_dim: int, seed: int = 7, uncond: bool = False):
    """'Cached T5': context ~ N(0,1), mask[i,:n_i] = True with n_i = 8 + (5 i mod 24); the
    unconditional row keeps only the first token (mirrors "" -> EOS)."""
    ctx = torch.randn(B, Lc, ctx_dim, generator=torch.Generator().manual_seed(seed))
    mask = torch.zeros(B, Lc, dtype=torch.bool)
    for i in range(B):
        n = 1 if uncond else min(Lc, 8 + (5 * i) % 24)
        mask[i, :n] = True
    return ctx, mask


def synth_gt(B: int, L: int, C: int = 128, seed: int = 11, lo: float = 0.25, hi: float = 0.75):
    """Inpainting stand-in: gt latent ~ N(0,1), mask True on frames [lo L, hi L) (edited span)."""
    gt = torch.randn(B, C, L, generator=torch.Generator().manual_seed(seed))
    m = torch.zeros(B, C, L, dtype=torch.bool)
    m[:, :, int(lo * L):int(hi * L)] = True
    return gt, m


def synth_energy_audio(B, T, seed=9):
    """Clips of very different loudness (1e-3 .. 0.3), slow amplitude modulation, a silent stretch in clip 0 (exercises the min_db floor)."""
    g = torch.Generator().manual_seed(seed)
    audio = torch.randn(B, T, generator=g) * torch.logspace(-3, -0.5, B)[:, None]
    audio[0, T // 3: T // 2] = 0.0
    return audio * (0.5 + 0.5 * torch.sin(torch.arange(T) / 2400.0))[None]


# transformers.T5Config fields of the text encoders the shipped configs name (ckpts/ezaudio-xl.yml / ezaudio-l.yml: google/flan-t5-xl, -large)
T5_XL = dict(vocab_size=32128, d_model=2048, d_kv=64, num_heads=32, d_ff=5120, num_layers=24, relative_attention_num_buckets=32,
             relative_attention_max_distance=128, layer_norm_epsilon=1e-6)
T5_LARGE = dict(vocab_size=32128, d_model=1024, d_kv=64, num_heads=16, d_ff=2816, num_layers=24, relative_attention_num_buckets=32,
                relative_attention_max_distance=128, layer_norm_epsilon=1e-6)


def tiny_t5(d_kv=64, heads=4, layers=2):
    return dict(vocab_size=512, d_model=192, d_kv=d_kv, num_heads=heads, d_ff=320, num_layers=layers, relative_attention_num_buckets=32,
                relative_attention_max_distance=128, layer_norm_epsilon=1e-6)


def synth_tokens(B, L, vocab, seed=11):
    """Token ids + attention mask like tokenizer(..., padding='max_length'): row i keeps n_i = 5 + 7 i (capped) tokens, then pad id 0;
    the last row is the empty prompt (EOS only)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(2, vocab, (B, L), generator=g)
    mask = torch.zeros(B, L, dtype=torch.long)
    for i in range(B):
        n = 1 if (i == B - 1 and B > 1) else min(L, 5 + 7 * i)
        mask[i, :n] = 1
        ids[i, n - 1] = 1  # EOS
        ids[i, n:] = 0
    return ids, mask
