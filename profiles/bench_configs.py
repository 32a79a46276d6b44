"""SURVEY.md §8(d) configurations C4 (XL + energy ControlNet, 8 prompts, CFG 3.5) and C5 (XL inpainting, 30 s, 100 steps, 2 prompts/GPU, CFG 3.5,
+ VAE encode/decode) through the public API objects, CUDA-event timed, with the algorithmic-FLOP roofline of SURVEY §8(d).
Not a bench.py line (bench.py measures BASELINE.json's metric on C2/C3) - a record for profiles/.

  python profiles/bench_configs.py [--configs C4,C5] [--reps 2]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ezaudio_b200 import api, synth  # noqa: E402
from ezaudio_b200.inference import inference  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="C4,C5")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
peak = 1426.5
try:
    pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    peak = float(pk.get("bf16_tflops_sustained", peak))
except Exception:
    pass


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


res = {}
if "C4" in a.configs:
    B = 8
    from ezaudio_b200 import config
    # SURVEY's C4 puts the energy ControlNet on the XL denoiser (the shipped ckpts/controlnet/energy_l.yml is the L model)
    params = dict(config.BUILTIN_CONTROLNET["energy"], model_name="EzAudio-XL", model=synth.XL_MODEL,
                  text_encoder=dict(model="google/flan-t5-xl", max_length=100, cfg=0.1))
    cn = api.EzAudio_ControlNet("energy", ckpt_path="synthetic:2", controlnet_path="synthetic:3", vae_path="synthetic:6", device=dev,
                                text_encoder=api.SyntheticTextEncoder(2048, 100), max_batch=B, params=params)
    wave = (0.1 * torch.randn(240000, generator=torch.Generator().manual_seed(9))).numpy()
    prompts = [f"synthetic prompt {i}" for i in range(B)]
    ms, out = timed(lambda: cn.generate_audio(prompts, wave, guidance_scale=3.5, guidance_rescale=0, ddim_steps=50, eta=1, conditioning_scale=1,
                                              random_seed=2024), a.reps)
    assert len(out[1]) == B and all(np.isfinite(w).all() for w in out[1])
    tflop = (50 * 16 * 1167.8 + 8 * 499.4) / 1e3  # the API never encodes the reference clip beyond its shape: VAE encode not counted
    res["C4"] = dict(workload="XL + energy ControlNet, 50 steps, 8 prompts, CFG 3.5 (effective batch 16), 10 s, via EzAudio_ControlNet.generate_audio (host in/out)",
                     ms_per_job=ms, audio_s_per_s=80 / (ms * 1e-3), algorithmic_tflop=tflop, tensor_roofline_frac=tflop / (ms * 1e-3) / peak)
    del cn
    torch.cuda.empty_cache()
if "C5" in a.configs:
    B, L = 2, 1500
    ez = api.EzAudio("s3_xl", ckpt_path="synthetic:2", vae_path="synthetic:6", device=dev, text_encoder=api.SyntheticTextEncoder(2048, 100), max_batch=B,
                     max_length_s=30)
    audio = 0.1 * torch.randn(B, 1, 480 * L, generator=torch.Generator().manual_seed(9)).to(dev)
    prompts = [f"synthetic prompt {i}" for i in range(B)]
    embeds = ez._text_embeds(prompts, [""])

    def job():
        gt = ez.autoencoder(audio=audio)
        mask = torch.zeros(B, 128, L, device=dev, dtype=torch.bool)
        mask[:, :, 250:1250] = True
        return inference(ez.autoencoder, ez.unet, gt, mask, None, None, ez.params, ez.noise_scheduler, prompts, None, L, 3.5, 0.0, 100, 1, 2024, dev,
                         text_embeds=embeds).cpu()

    ms, out = timed(job, a.reps)
    assert out.shape == (B, 1, 480 * L) and torch.isfinite(out).all()
    tflop = (100 * 4 * 2528 + 2 * 3 * (499.4 + 499.3)) / 1e3
    res["C5"] = dict(workload="XL inpainting, 30 s (L=1500), 100 steps, 2 prompts on 1 GPU, CFG 3.5 (effective batch 4), VAE encode + decode, host waveform out",
                     ms_per_job=ms, audio_s_per_s=60 / (ms * 1e-3), algorithmic_tflop=tflop, tensor_roofline_frac=tflop / (ms * 1e-3) / peak)
if "T5" in a.configs:
    # the step before the path: flan-T5-XL encoder, 4 prompts + the empty negative prompt, 100 tokens each
    from ezaudio_b200 import weights
    from ezaudio_b200.t5 import T5EncoderModel
    cfg = synth.T5_XL
    sd = weights.synthetic_state_dict(weights.t5_param_shapes(cfg), 15)
    t5 = T5EncoderModel(cfg, max_batch=5, max_len=100, device=dev).load_state_dict(sd)
    del sd
    ids, mask = synth.synth_tokens(5, 100, cfg["vocab_size"])
    ids, mask = ids.to(dev), mask.to(dev)
    ms, out = timed(lambda: t5(input_ids=ids, attention_mask=mask).last_hidden_state, 5)
    assert out.shape == (5, 100, 2048) and torch.isfinite(out).all()
    wbytes = 2 * 24 * (3 * 2048 * 2048 + 2048 * 2048 + 3 * 5120 * 2048)
    gflop = 2 * 500 * 24 * (4 * 2048 * 2048 + 3 * 5120 * 2048) / 1e9
    res["T5"] = dict(workload="flan-T5-XL encoder (24 layers, d_model 2048), 5 prompts x 100 tokens, bf16 operands", ms_per_encode=ms,
                     weight_bytes=wbytes, hbm_GBps=wbytes / (ms * 1e-3) / 1e9, gflop=gflop, note="weight-bandwidth bound: every linear reads its bf16 weight once for 500 rows")
print(json.dumps(res))
