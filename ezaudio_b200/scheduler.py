"""DDIM scheduler (host-side control logic) -- restatement of `diffusers.DDIMScheduler` for the configuration the
reference instantiates (api/ezaudio.py:92-97 with ckpts/ezaudio-xl.yml:52-60; call sites src/inference.py:64,71,98-100).
diffusers is a third-party, un-pinned, un-vendored dependency of the reference (requirements.txt:2) and is absent from this
image, so the algorithm is restated from its published form; parity at this boundary is unpinned
upstream and is checked by closed-form invariants (tests/test_scheduler.py).

Only the scalar schedule lives here; the tensor update runs in the fused CUDA kernel `ezb_cfg_ddim_step`.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch


class DDIMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 prediction_type="v_prediction", rescale_betas_zero_snr=True, timestep_spacing="trailing", clip_sample=False,
                 set_alpha_to_one=True, **unused):
        if beta_schedule != "scaled_linear" or prediction_type != "v_prediction" or clip_sample or timestep_spacing != "trailing":
            raise NotImplementedError("only the shipped diffusion config (scaled_linear, v_prediction, trailing, no clipping)")
        self.num_train_timesteps = num_train_timesteps
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        if rescale_betas_zero_snr:
            abar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
            s0, sT = abar_sqrt[0].clone(), abar_sqrt[-1].clone()
            abar_sqrt = (abar_sqrt - sT) * (s0 / (s0 - sT))
            abar = abar_sqrt ** 2
            alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
            betas = 1 - alphas
        self.betas = betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps: int):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps exceeds num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps / num_inference_steps
        ts = np.round(np.arange(self.num_train_timesteps, 0, -step_ratio)).astype(np.int64) - 1
        self.timesteps = torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step_coefficients(self, timestep: int, eta: float) -> List[float]:
        """[sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma] of DDIMScheduler.step (fp32, diffusers op order):
        x0 = sqrt(a) x - sqrt(1-a) v ; eps = sqrt(a) v + sqrt(1-a) x ; prev = sqrt(a_prev) x0 + sqrt(1-a_prev-s^2) eps + s z."""
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a = self.alphas_cumprod[t]
        ap = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b = 1 - a
        variance = ((1 - ap) / b) * (1 - a / ap)
        sigma = eta * variance ** 0.5
        dirc = (1 - ap - sigma ** 2).clamp_min(0) ** 0.5  # radicand is >= 0 for the shipped schedule (tested); clamp guards round-off
        return [float(a ** 0.5), float(b ** 0.5), float(ap ** 0.5), float(dirc), float(sigma)]
