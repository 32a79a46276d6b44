"""Weight wire format of the reference and synthetic weights.

`dit_param_shapes` / `controlnet_param_shapes` / `vae_decoder_param_shapes` enumerate the exact
state-dict keys and shapes the reference modules produce
(`MaskDiT(**cfg).state_dict()`, src/models/conditioners.py:123-135 + src/models/udit.py:11-180;
`DiTControlNet`, src/models/controlnet.py:87-236; `OobleckDecoder`,
src/modules/stable_vae/models/autoencoders.py:149-187).  They are what the C-ABI loader
validates incoming checkpoints against; tests check them against the live reference.

`synthetic_state_dict` draws a deterministic random checkpoint in that format.  There is no
network and no shipped checkpoint, and the reference's own init zeroes half the
block, so benchmarks and parity tests use these weights everywhere (the same bits
are loaded into the reference modules when goldens are generated).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

Shapes = "OrderedDict[str, Tuple[int, ...]]"


def check_dit_config(cfg: dict) -> None:
    """Only the shipped combination has a kernel path (ckpts/ezaudio-xl.yml:5-37); anything
    else raises instead of silently falling back."""
    want = dict(input_type="1d", patch_size=1, qkv_bias=False, qk_norm="layernorm",
                norm_layer="layernorm", act_layer="geglu", context_norm=True,
                time_fusion="ada_sola_bias", context_fusion="cross", pe_method="none",
                context_pe_method="none", rope_mode="shared", use_conv=True, skip=True,
                skip_norm=True, cls_dim=None, qk_scale=None)
    for k, v in want.items():
        if cfg.get(k, v) != v:
            raise NotImplementedError(f"ezaudio_b200: unsupported DiT config {k}={cfg.get(k)!r} (only {v!r})")
    D, H = cfg["embed_dim"], cfg["num_heads"]
    if D % H or (D // H) % 8 or D % 8:
        raise NotImplementedError(f"embed_dim={D}, num_heads={H}: head_dim must be a multiple of 8")
    if cfg["depth"] % 2:
        raise NotImplementedError("depth must be even")


def _block_shapes(out: Dict, p: str, D: int, H: int, inner: int, r: int, skip: bool):
    dh = D // H
    for n in ("norm1", "norm2", "norm3", "norm_context"):
        out[f"{p}.{n}.weight"] = (D,)
        out[f"{p}.{n}.bias"] = (D,)
    for a in ("attn", "cross_attn"):
        for w in ("to_q", "to_k", "to_v"):
            out[f"{p}.{a}.{w}.weight"] = (D, D)
        for n in ("norm_q", "norm_k"):
            out[f"{p}.{a}.{n}.weight"] = (dh,)
            out[f"{p}.{a}.{n}.bias"] = (dh,)
        out[f"{p}.{a}.proj.weight"] = (D, D)
        out[f"{p}.{a}.proj.bias"] = (D,)
        if a == "attn":
            out[f"{p}.{a}.rotary.inv_freq"] = (dh // 2,)
    out[f"{p}.mlp.net.0.proj.weight"] = (2 * inner, D)
    out[f"{p}.mlp.net.0.proj.bias"] = (2 * inner,)
    out[f"{p}.mlp.net.2.weight"] = (D, inner)
    out[f"{p}.mlp.net.2.bias"] = (D,)
    out[f"{p}.adaln.scale_shift_table"] = (6, D)
    out[f"{p}.adaln.lora_a.weight"] = (6 * r, D)
    out[f"{p}.adaln.lora_b.weight"] = (6 * D, 6 * r)
    if skip:
        out[f"{p}.skip_norm.weight"] = (2 * D,)
        out[f"{p}.skip_norm.bias"] = (2 * D,)
        out[f"{p}.skip_linear.weight"] = (D, 2 * D)
        out[f"{p}.skip_linear.bias"] = (D,)


def _trunk_shapes(out: Dict, p: str, cfg: dict, final: bool):
    D, Cin, ctx = cfg["embed_dim"], cfg["in_chans"], cfg["context_dim"]
    out[f"{p}patch_embed.proj.weight"] = (D, Cin, 1)
    out[f"{p}patch_embed.proj.bias"] = (D,)
    out[f"{p}time_embed.mlp.0.weight"] = (D, 256)
    out[f"{p}time_embed.mlp.0.bias"] = (D,)
    out[f"{p}time_embed.mlp.2.weight"] = (D, D)
    out[f"{p}time_embed.mlp.2.bias"] = (D,)
    if final:
        out[f"{p}time_ada_final.weight"] = (2 * D, D)
        out[f"{p}time_ada_final.bias"] = (2 * D,)
    out[f"{p}time_ada.weight"] = (6 * D, D)
    out[f"{p}time_ada.bias"] = (6 * D,)
    out[f"{p}context_embed.0.weight"] = (D, ctx)
    out[f"{p}context_embed.0.bias"] = (D,)
    out[f"{p}context_embed.2.weight"] = (D, D)
    out[f"{p}context_embed.2.bias"] = (D,)


def dit_param_shapes(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    """Keys/shapes of `MaskDiT(**cfg).state_dict()`."""
    check_dit_config(cfg)
    D, H = cfg["embed_dim"], cfg["num_heads"]
    inner, r, half = int(D * cfg["mlp_ratio"]), cfg["ada_sola_rank"], cfg["depth"] // 2
    out: Dict = OrderedDict()
    out["mask_embed"] = (cfg["out_chans"],)
    _trunk_shapes(out, "model.", cfg, final=True)
    for i in range(half):
        _block_shapes(out, f"model.in_blocks.{i}", D, H, inner, r, False)
    _block_shapes(out, "model.mid_block", D, H, inner, r, False)
    for i in range(half):
        _block_shapes(out, f"model.out_blocks.{i}", D, H, inner, r, True)
    out["model.final_block.norm.weight"] = (D,)
    out["model.final_block.norm.bias"] = (D,)
    out["model.final_block.linear.weight"] = (cfg["out_chans"], D)
    out["model.final_block.linear.bias"] = (cfg["out_chans"],)
    out["model.final_block.final_layer.weight"] = (cfg["out_chans"], cfg["out_chans"], 3)
    out["model.final_block.final_layer.bias"] = (cfg["out_chans"],)
    return out


def controlnet_param_shapes(cfg: dict, cn: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    """Keys/shapes of `DiTControlNet(**cfg, **cn).state_dict()` (controlnet.py:87-236):
    trunk without time_ada_final / mid / out / final blocks, plus the stem and zero-linears."""
    check_dit_config(cfg)
    D, H = cfg["embed_dim"], cfg["num_heads"]
    inner, r, half = int(D * cfg["mlp_ratio"]), cfg["ada_sola_rank"], cfg["depth"] // 2
    out: Dict = OrderedDict()
    _trunk_shapes(out, "", cfg, final=False)
    for i in range(half):
        _block_shapes(out, f"in_blocks.{i}", D, H, inner, r, False)
    blocks = list(cn["cond_blocks"])
    out["controlnet_pre.conv_in.weight"] = (blocks[0], cn["cond_in"], 1)
    out["controlnet_pre.conv_in.bias"] = (blocks[0],)
    if cn.get("cond_mask", False):
        out["controlnet_pre.mask_embed"] = (blocks[0],)
        blocks[0] += 1  # controlnet.py:23
    for i in range(len(blocks) - 1):
        out[f"controlnet_pre.blocks.{i}.0.weight"] = (blocks[i], blocks[i], 3)
        out[f"controlnet_pre.blocks.{i}.0.bias"] = (blocks[i],)
        out[f"controlnet_pre.blocks.{i}.2.weight"] = (blocks[i + 1], blocks[i], 3)
        out[f"controlnet_pre.blocks.{i}.2.bias"] = (blocks[i + 1],)
    out["controlnet_pre.conv_out.weight"] = (D, blocks[-1], 1)
    out["controlnet_pre.conv_out.bias"] = (D,)
    for i in range(half):
        out[f"controlnet_zero_blocks.{i}.weight"] = (D, D)
        out[f"controlnet_zero_blocks.{i}.bias"] = (D,)
    return out


def vae_decoder_param_shapes(dec_cfg: dict, prefix: str = "decoder.") -> "OrderedDict[str, Tuple[int, ...]]":
    """Keys/shapes of the `decoder.*` slice of the VAE state-dict
    (stable_vae/models/autoencoders.py:149-187; ckpts/vae/config.json:18-28)."""
    ch, mults, strides = dec_cfg["channels"], [1] + list(dec_cfg["c_mults"]), list(dec_cfg["strides"])
    latent, outc = dec_cfg["latent_dim"], dec_cfg["out_channels"]
    if not dec_cfg.get("use_snake", False) or dec_cfg.get("final_tanh", True):
        raise NotImplementedError("ezaudio_b200 VAE decoder: only use_snake=true, final_tanh=false")
    out: Dict = OrderedDict()
    p = prefix + "layers."

    def wn(key, co, ci, k, bias=True, transpose=False):
        out[key + ".weight_g"] = ((ci if transpose else co), 1, 1)
        out[key + ".weight_v"] = (ci, co, k) if transpose else (co, ci, k)
        if bias:
            out[key + ".bias"] = (co,)

    def snake(key, c):
        out[key + ".alpha"] = (c,)
        out[key + ".beta"] = (c,)

    wn(p + "0", mults[-1] * ch, latent, 7)
    j = 1
    for i in range(len(mults) - 1, 0, -1):
        cin, cout, s = mults[i] * ch, mults[i - 1] * ch, strides[i - 1]
        q = f"{p}{j}.layers."
        snake(q + "0", cin)
        wn(q + "1", cout, cin, 2 * s, transpose=True)
        for u in range(3):
            snake(f"{q}{u + 2}.layers.0", cout)
            wn(f"{q}{u + 2}.layers.1", cout, cout, 7)
            snake(f"{q}{u + 2}.layers.2", cout)
            wn(f"{q}{u + 2}.layers.3", cout, cout, 1)
        j += 1
    snake(f"{p}{j}", mults[0] * ch)
    wn(f"{p}{j + 1}", outc, mults[0] * ch, 7, bias=False)
    return out


def vae_encoder_param_shapes(enc_cfg: dict, prefix: str = "encoder.") -> "OrderedDict[str, Tuple[int, ...]]":
    """Keys/shapes of the `encoder.*` slice of the VAE state-dict (stable_vae/models/autoencoders.py:115-146;
    ckpts/vae/config.json:7-16: in 1 ch, channels 128, c_mults [1,2,4,8], strides [2,4,6,10], latent 256 = mean|scale)."""
    ch, mults, strides = enc_cfg["channels"], [1] + list(enc_cfg["c_mults"]), list(enc_cfg["strides"])
    if not enc_cfg.get("use_snake", False):
        raise NotImplementedError("ezaudio_b200 VAE encoder: only use_snake=true")
    out: Dict = OrderedDict()
    p = prefix + "layers."

    def wn(key, co, ci, k):
        out[key + ".weight_g"] = (co, 1, 1)
        out[key + ".weight_v"] = (co, ci, k)
        out[key + ".bias"] = (co,)

    def snake(key, c):
        out[key + ".alpha"] = (c,)
        out[key + ".beta"] = (c,)

    wn(p + "0", mults[0] * ch, enc_cfg["in_channels"], 7)
    for i in range(len(mults) - 1):
        cin, cout, s = mults[i] * ch, mults[i + 1] * ch, strides[i]
        q = f"{p}{i + 1}.layers."
        for u in range(3):
            snake(f"{q}{u}.layers.0", cin)
            wn(f"{q}{u}.layers.1", cin, cin, 7)
            snake(f"{q}{u}.layers.2", cin)
            wn(f"{q}{u}.layers.3", cin, cin, 1)
        snake(q + "3", cin)
        wn(q + "4", cout, cin, 2 * s)
    n = len(mults)
    snake(f"{p}{n}", mults[-1] * ch)
    wn(f"{p}{n + 1}", enc_cfg["latent_dim"], mults[-1] * ch, 3)
    return out


def t5_param_shapes(cfg) -> "OrderedDict[str, tuple]":
    """State-dict keys / shapes of transformers.T5EncoderModel (the text encoder the reference loads, api/ezaudio.py:78-79) for a
    gated-GELU T5 v1.1 / flan-T5 config dict: vocab_size, d_model, d_kv, num_heads, d_ff, num_layers, relative_attention_num_buckets."""
    D, inner, F = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"]
    out = OrderedDict()
    out["shared.weight"] = (cfg["vocab_size"], D)
    for i in range(cfg["num_layers"]):
        a = f"encoder.block.{i}.layer.0."
        for n in ("q", "k", "v"):
            out[a + f"SelfAttention.{n}.weight"] = (inner, D)
        out[a + "SelfAttention.o.weight"] = (D, inner)
        if i == 0:
            out[a + "SelfAttention.relative_attention_bias.weight"] = (cfg["relative_attention_num_buckets"], cfg["num_heads"])
        out[a + "layer_norm.weight"] = (D,)
        f = f"encoder.block.{i}.layer.1."
        out[f + "DenseReluDense.wi_0.weight"] = (F, D)
        out[f + "DenseReluDense.wi_1.weight"] = (F, D)
        out[f + "DenseReluDense.wo.weight"] = (D, F)
        out[f + "layer_norm.weight"] = (D,)
    out["encoder.final_layer_norm.weight"] = (D,)
    return out


def synthetic_state_dict(shapes, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic random checkpoint (CPU generator; identical bits wherever the same torch
    build runs).  Every tensor the reference zero-initialises is drawn non-zero,
    scaled so activations stay O(1) through the full depth."""
    out: Dict[str, torch.Tensor] = OrderedDict()
    for idx, (k, shp) in enumerate(shapes.items()):
        g = torch.Generator().manual_seed(seed * 1_000_003 + idx)
        rn = lambda *s, std=1.0: torch.randn(*s, generator=g, dtype=torch.float32) * std
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "inv_freq":  # rotary.py:41-42
            dh = shp[0] * 2
            t = 1.0 / (10000 ** (torch.arange(0, dh, 2).float() / dh))
        elif leaf in ("alpha", "beta"):  # SnakeBeta log-scale params
            t = rn(*shp, std=0.3)
        elif leaf == "weight_g":
            t = None  # filled after weight_v below
        elif leaf == "weight_v":
            fan_in = shp[1] * shp[2]
            t = rn(*shp, std=0.6 / math.sqrt(fan_in))
        elif k == "mask_embed" or k.endswith("controlnet_pre.mask_embed"):
            t = rn(*shp, std=0.5)
        elif leaf == "scale_shift_table":
            t = rn(*shp, std=0.1)
        elif leaf == "bias":
            t = rn(*shp, std=0.05 if ("norm" in k) else 0.02)
        elif leaf == "weight" and len(shp) == 1:  # LayerNorm gains
            t = 1.0 + rn(*shp, std=0.1)
        elif leaf == "weight":
            fan_in = math.prod(shp[1:])
            gain = 1.0
            if "time_ada" in k or "lora_b" in k:
                gain = 0.3
            if "controlnet_zero_blocks" in k or "conv_out" in k:
                gain = 0.5
            if ".SelfAttention.q." in k:  # T5 folds the 1/sqrt(d_kv) of the (unscaled) attention into the query init
                gain = 0.125
            t = rn(*shp, std=gain / math.sqrt(fan_in))
        else:
            raise KeyError(k)
        out[k] = t
    for k in list(out):
        if k.endswith("weight_g"):
            v = out[k[:-1] + "v"]
            g = torch.Generator().manual_seed(seed * 1_000_003 + 7_777_777 + len(k))
            n = v.flatten(1).norm(dim=1).view(-1, 1, 1)
            out[k] = n * (1.0 + 0.1 * torch.randn(n.shape, generator=g))
    return {k: v.to(dtype).contiguous() for k, v in out.items()}
