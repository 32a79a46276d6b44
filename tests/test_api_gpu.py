"""Public API conformance on the GPU (api/ezaudio.py:101-130, api/controlnet.py:113-161): signatures, return types, shapes,
determinism under a fixed seed, batched extension.  Uses the tiny architecture through `config_path`-free params injection."""
import inspect

import numpy as np
import pytest
import torch

from ezaudio_b200 import config, synth
from tests import helpers

pytestmark = pytest.mark.gpu


def _tiny_params():
    p = config.load_params("s3_xl")
    p["model"] = synth.tiny_model(72)
    p["text_encoder"] = dict(p["text_encoder"], max_length=16)
    return p


def test_signatures_match_reference():
    from ezaudio_b200.api import EzAudio, EzAudio_ControlNet
    g = inspect.signature(EzAudio.generate_audio).parameters
    assert list(g)[:9] == ["self", "text", "length", "guidance_scale", "guidance_rescale", "ddim_steps", "eta", "random_seed", "randomize_seed"]
    assert (g["length"].default, g["guidance_scale"].default, g["guidance_rescale"].default, g["ddim_steps"].default, g["eta"].default) == (10, 5, 0.75, 100, 1)
    e = inspect.signature(EzAudio.editing_audio).parameters
    assert list(e)[:6] == ["self", "text", "boundary", "gt_file", "mask_start", "mask_length"]
    assert (e["guidance_scale"].default, e["guidance_rescale"].default, e["ddim_steps"].default) == (3.5, 0, 100)
    c = inspect.signature(EzAudio_ControlNet.generate_audio).parameters
    assert list(c)[:4] == ["self", "text", "audio_path", "surpass_noise"]
    assert (c["guidance_scale"].default, c["ddim_steps"].default, c["conditioning_scale"].default) == (3.5, 50, 1)


def test_generate_audio_shapes_types_determinism(monkeypatch):
    from ezaudio_b200 import api
    tiny = _tiny_params()
    monkeypatch.setattr(config, "load_params", lambda name, path=None, table=None: tiny)
    enc = api.SyntheticTextEncoder(64, 16)
    ez = api.EzAudio("s3_xl", ckpt_path="synthetic:3", vae_path="synthetic:6", text_encoder=enc, max_batch=2, max_length_s=2,
                     vae_config_path=None)
    sr, wav = ez.generate_audio("a dog barks", length=1, ddim_steps=4, random_seed=7)
    assert sr == 24000 and isinstance(wav, np.ndarray) and wav.dtype == np.float32 and wav.shape == (24000,)
    assert np.isfinite(wav).all()
    sr2, wav2 = ez.generate_audio("a dog barks", length=1, ddim_steps=4, random_seed=7)
    assert np.array_equal(wav, wav2)                      # same seed -> same audio (graph replay path on the 2nd call)
    _, wav3 = ez.generate_audio("a dog barks", length=1, ddim_steps=4, random_seed=8)
    assert not np.array_equal(wav, wav3)
    sr, batch = ez.generate_audio(["a dog barks", "rain on a roof"], length=1, ddim_steps=4, random_seed=7)
    assert isinstance(batch, list) and len(batch) == 2 and batch[0].shape == (24000,)
    assert np.allclose(batch[0], wav, atol=2e-2)          # prompt 0 of a batch == the single-prompt run (bf16 batch-size noise only)
    _, nocfg = ez.generate_audio("", length=1, ddim_steps=4, random_seed=7)   # text == '' -> no CFG branch (api/ezaudio.py:109-111)
    assert nocfg.shape == (24000,) and np.isfinite(nocfg).all()


def test_controlnet_generate_audio(monkeypatch):
    from ezaudio_b200 import api
    p = _tiny_params()
    p["controlnet"] = synth.CONTROLNET
    p["conditioner"] = config.BUILTIN_CONTROLNET["energy"]["conditioner"]
    enc = api.SyntheticTextEncoder(64, 16)
    ez = api.EzAudio_ControlNet("energy", ckpt_path="synthetic:5", controlnet_path="synthetic:6", vae_path="synthetic:6", text_encoder=enc, max_batch=1,
                                params=p)
    ref_audio = (0.1 * torch.randn(3 * 24000, generator=torch.Generator().manual_seed(9))).numpy()
    sr, wav = ez.generate_audio("a siren", ref_audio, ddim_steps=3, random_seed=1)
    assert sr == 24000 and wav.dtype == np.float32 and wav.shape == (3 * 24000,) and np.isfinite(wav).all()


def test_editing_audio_end_to_end(monkeypatch, tmp_path):
    """api/ezaudio.py:132-207: crop -> VAE encode (stochastic bottleneck) -> masked sampling -> paste -> decode -> splice."""
    from scipy.io import wavfile
    from ezaudio_b200 import api
    tiny = _tiny_params()
    monkeypatch.setattr(config, "load_params", lambda name, path=None, table=None: tiny)
    ez = api.EzAudio("s3_xl", ckpt_path="synthetic:3", vae_path="synthetic:6", text_encoder=api.SyntheticTextEncoder(64, 16), max_batch=1,
                     max_length_s=6)
    sr = 24000
    t = np.arange(4 * sr) / sr
    wav = (0.3 * np.sin(2 * np.pi * 220 * t)).astype(np.float32)
    f = str(tmp_path / "in.wav")
    wavfile.write(f, sr, (wav * 32767).astype(np.int16))
    torch.manual_seed(0)
    out_sr, out = ez.editing_audio("a bell", boundary=1, gt_file=f, mask_start=1.5, mask_length=1.0, ddim_steps=3, random_seed=3)
    assert out_sr == sr and out.dtype == np.float32 and out.shape == (4 * sr,) and np.isfinite(out).all()
    ref = wav / (np.abs(wav).max() + 1e-9)
    assert np.allclose(out[: int(0.4 * sr)], ref[: int(0.4 * sr)], atol=1e-3)     # outside [mask-boundary, mask+boundary]: untouched original
    assert not np.allclose(out[int(1.6 * sr): int(2.4 * sr)], ref[int(1.6 * sr): int(2.4 * sr)], atol=1e-2)  # edited span regenerated


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("energy_api", dict(hop_size=240, window_size=1920, min_db=-60.0, norm=True)),
                                     ("energy_quant", dict(hop_size=512, window_size=1024, min_db=-80.0, norm=True, quantize_levels=16))])
def test_energy_condition_kernel_matches_reference(name, kw):
    """ezb_energy_condition vs the UNMODIFIED reference EnergyExtractor's golden output (and the oracle on a 10-s clip)."""
    from ezaudio_b200.api import energy_condition
    from oracle import ezaudio_oracle as O
    g = helpers.load_golden(name)
    audio = synth.synth_energy_audio(int(g["B"]), int(g["T"]), int(g["seed"]))
    out = energy_condition(audio.cuda(), **kw).cpu()
    assert out.shape == (int(g["B"]), 1, int(g["T"]) // kw["hop_size"])
    assert float((out[:, 0] - torch.from_numpy(g["out"])[..., 0]).abs().max()) < 2e-5
    long = synth.synth_energy_audio(4, 240000, 3)
    got = energy_condition(long.cuda(), hop_size=240, window_size=1920, min_db=-60, norm=True).cpu()
    want = O.energy_extract(long, 240, 1920, -60.0, True)
    assert got.shape == (4, 1, 1000) and float((got[:, 0] - want[..., 0]).abs().max()) < 2e-5


def test_generate_audio_from_text_with_native_t5(monkeypatch):
    """`generate_audio(str)` with nothing but native kernels between the string and the waveform: tokenizer stand-in -> T5 encoder
    (ezb_t5_forward) -> DiT loop -> VAE decode.  The T5 width must match the denoiser's context_dim (udit.py:94)."""
    from ezaudio_b200 import api, weights
    from ezaudio_b200.t5 import T5EncoderModel
    tiny = _tiny_params()
    monkeypatch.setattr(config, "load_params", lambda name, path=None, table=None: tiny)
    tcfg = dict(synth.tiny_t5(), d_model=tiny["model"]["context_dim"])
    t5 = T5EncoderModel(tcfg, max_batch=4, max_len=16).load_state_dict(weights.synthetic_state_dict(weights.t5_param_shapes(tcfg), 12))
    enc = api.NativeTextEncoder(api.HashTokenizer(tcfg["vocab_size"]), t5, 16)
    e, m = enc(["a dog barks twice", ""])
    assert e.shape == (2, 16, tcfg["d_model"]) and m.sum(1).tolist() == [5, 1]
    ez = api.EzAudio("s3_xl", ckpt_path="synthetic:3", vae_path="synthetic:6", text_encoder=enc, max_batch=2, max_length_s=2)
    sr, wav = ez.generate_audio("a dog barks twice", length=1, ddim_steps=4, random_seed=7)
    assert sr == 24000 and wav.shape == (24000,) and np.isfinite(wav).all()
    _, other = ez.generate_audio("heavy rain on a tin roof", length=1, ddim_steps=4, random_seed=7)
    assert not np.array_equal(wav, other)   # the text actually conditions the result
