"""CUDA Oobleck decoder vs golden outputs of the UNMODIFIED reference decoder (fp32 CPU).
Tolerance is relative to the output scale (random-init outputs have |max| ~ 0.1):
  bf16x3 parity mode: max-abs < 1e-3 * max|ref| + 1e-5 ;  bf16 fast mode: < 6e-2 * max|ref|  -- the reference's own
  bf16-autocast decoder differs from its fp32 output by ~5 % of |max| (1.6e-3 on |max| 0.03)."""
import pytest
import torch

from ezaudio_b200 import synth, weights
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,rel", [("bf16x3", 1e-3), ("bf16", 6e-2)])
@pytest.mark.parametrize("name,dcfg,B,L", [("vae_tiny", synth.tiny_vae(16), 2, 9), ("vae_full", synth.VAE_DECODER, 1, 12)])
def test_vae_decode_matches_reference(name, dcfg, B, L, precision, rel):
    from ezaudio_b200.vae import OobleckDecoder
    g = helpers.load_golden(name)
    sd = weights.synthetic_state_dict(weights.vae_decoder_param_shapes(dcfg), 6)
    dec = OobleckDecoder(precision=precision, max_batch=B, max_latent_len=L, **dcfg).load_state_dict(sd)
    z = synth.synth_latents(B, L, dcfg["latent_dim"], seed=31).cuda()
    wav = dec(z)
    torch.cuda.synchronize()
    ref = torch.from_numpy(g["out"])
    assert wav.shape == ref.shape
    err = float((wav.cpu() - ref).abs().max())
    assert err < rel * float(ref.abs().max()) + 1e-5, (err, float(ref.abs().max()))
