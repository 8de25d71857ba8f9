"""FlowMatchEulerDiscreteScheduler (host side) with the FLUX.1-dev scheduler_config.json values.

Restates the diffusers scheduler the reference drives at train_flux/flux/generate.py:193-213,276:
sigmas are produced in float32 with numpy exactly as diffusers does, so the schedule is
bit-identical; the update itself runs on the GPU (rf_op_euler_step, or fused in rf_dit_denoise)."""
from __future__ import annotations

import math

import numpy as np
import torch


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def calculate_shift(image_seq_len, base_seq_len: int = 256, max_seq_len: int = 4096,
                    base_shift: float = 0.5, max_shift: float = 1.15):
    """diffusers.pipelines.flux.pipeline_flux.calculate_shift (generate.py:195-201)."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


class FlowMatchEulerDiscreteScheduler:
    order = 1

    def __init__(self, **overrides):
        self.config = _Cfg(num_train_timesteps=1000, shift=3.0, use_dynamic_shifting=True,
                           base_shift=0.5, max_shift=1.15, base_image_seq_len=256,
                           max_image_seq_len=4096)
        self.config.update(overrides)
        self.timesteps = None
        self.sigmas = None
        self._step_index = None

    def time_shift(self, mu: float, sigma: float, t):
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None):
        if sigmas is None:
            sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
        sigmas = np.array(sigmas).astype(np.float32)
        if self.config.use_dynamic_shifting:
            if mu is None:
                raise ValueError("mu is required with use_dynamic_shifting")
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            s = self.config.shift
            sigmas = s * sigmas / (1 + (s - 1) * sigmas)
        sigmas = torch.from_numpy(np.asarray(sigmas, dtype=np.float32)).to(torch.float32)
        self.timesteps = (sigmas * self.config.num_train_timesteps).to(device or "cpu")
        self.sigmas = torch.cat([sigmas, torch.zeros(1, dtype=torch.float32)])
        self._step_index = None

    @property
    def step_index(self):
        return self._step_index

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict=True):
        """x_{i+1} = (float(x_i) + (sigma_{i+1} - sigma_i) * v).to(v.dtype) on the GPU."""
        from . import _lib as L
        if self._step_index is None:
            self._step_index = 0
        dev = sample.device
        x = sample.to(torch.bfloat16).contiguous().clone()
        v = model_output.to(torch.bfloat16).contiguous()
        sig = self.sigmas.to(dev)
        idx = torch.tensor([self._step_index], dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.load().rf_op_euler_step(L.ptr(x), L.ptr(v), L.ptr(sig), L.ptr(idx), x.numel(),
                                              L.cur_stream()), "rf_op_euler_step")
        self._step_index += 1
        return (x,)
