"""Noise schedule tables (host, float64) and the per-step coefficient rows the sampler kernels index on device.

Mirrors DDPMScheduler (LM:1241-1276) with the "cosine" betas of LM:1145-1162,1217-1221, plus the respacing rule of
diffusion/respace.py:73-87 for strided samplers.  The reference looks coefficients up with
``torch.from_numpy(arr)[t].float()`` (LM:1235): float64 table entry rounded to fp32 — ``coef_rows`` does the same.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np


class DDPMScheduler:
    def __init__(self, timesteps: int = 200, scale: float = 1.0, *, betas: np.ndarray = None):
        """Same positional signature as the reference (LM:1242: timesteps, scale); `betas` (keyword only) builds a
        re-spaced schedule."""
        self.scale = scale
        if betas is None:
            ab = lambda u: math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2
            betas = np.array([min(1 - ab((i + 1) / timesteps) / ab(i / timesteps), 0.999) for i in range(timesteps)],
                             dtype=np.float64)
        self.num_timesteps = len(betas)
        self.betas = betas
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)

    def spaced(self, use_timesteps: Sequence[int]) -> Tuple["DDPMScheduler", List[int]]:
        """respace.py:73-87: betas re-derived over the kept steps; returns (scheduler, timestep_map)."""
        keep = sorted(set(int(t) for t in use_timesteps))
        last, nb = 1.0, []
        for i, ac in enumerate(self.alphas_cumprod):
            if i in keep:
                nb.append(1 - ac / last)
                last = ac
        return DDPMScheduler(betas=np.array(nb, dtype=np.float64)), keep

    # ---- the reference's lookups (LM:1226-1238, :1278-1297): float64 table entry -> fp32 -> broadcast to `shape`
    @staticmethod
    def _extract(arr: np.ndarray, t, shape):
        import torch
        res = torch.from_numpy(arr).to(device=t.device)[t].float()
        while res.dim() < len(shape):
            res = res[..., None]
        return res.expand(shape)

    def get_beta(self, t, shape):
        return self._extract(self.betas, t, shape)

    def get_sqrt_alpha_cum(self, t, shape):
        return self._extract(self.sqrt_alphas_cumprod, t, shape)

    def get_sqrt_one_minus_alpha_cum(self, t, shape):
        return self._extract(self.sqrt_one_minus_alphas_cumprod, t, shape)

    def get_alpha_cum(self, t, shape):
        return self._extract(self.alphas_cumprod, t, shape)

    def get_alpha_prev_cum(self, t, shape):
        return self._extract(self.alphas_cumprod_prev, t, shape)

    def get_snr(self, t):
        return self.get_sqrt_alpha_cum(t, t.shape) ** 2 / self.get_sqrt_one_minus_alpha_cum(t, t.shape) ** 2

    # ---- device coefficient rows (float32 x 8 per step) -------------------------------------------------------
    def ddim_rows(self) -> np.ndarray:
        """[0] sqrt_ab [1] sqrt(1-ab) [2] sqrt(ab_prev) [3] sqrt(1-ab_prev) [4] sqrt(1/ab) [5] sqrt(1/ab-1)."""
        f = lambda a: a.astype(np.float32)
        abp = f(self.alphas_cumprod_prev)  # the reference rounds ab_prev to fp32 first, then takes sqrt in fp32
        rows = np.zeros((self.num_timesteps, 8), dtype=np.float32)
        rows[:, 0] = f(self.sqrt_alphas_cumprod)
        rows[:, 1] = f(self.sqrt_one_minus_alphas_cumprod)
        rows[:, 2] = np.sqrt(abp)
        rows[:, 3] = np.sqrt(np.float32(1.0) - abp)
        rows[:, 4] = f(self.sqrt_recip_alphas_cumprod)
        rows[:, 5] = f(self.sqrt_recipm1_alphas_cumprod)
        return rows

    def ddpm_rows(self, large_var: bool = False) -> np.ndarray:
        """[0] sqrt(1/ab) [1] sqrt(1/ab-1) [2] coef1 [3] coef2 [4] exp(.5 logvar) (0 at t = 0)."""
        f = lambda a: a.astype(np.float32)
        if large_var:  # gaussian_diffusion.py:295-303 FIXED_LARGE
            logvar = np.log(np.append(self.posterior_variance[1], self.betas[1:]))
        else:
            logvar = self.posterior_log_variance_clipped
        rows = np.zeros((self.num_timesteps, 8), dtype=np.float32)
        rows[:, 0] = f(self.sqrt_recip_alphas_cumprod)
        rows[:, 1] = f(self.sqrt_recipm1_alphas_cumprod)
        rows[:, 2] = f(self.posterior_mean_coef1)
        rows[:, 3] = f(self.posterior_mean_coef2)
        rows[:, 4] = np.exp(np.float32(0.5) * f(logvar))
        rows[0, 4] = 0.0  # nonzero_mask (gaussian_diffusion.py:411-413)
        return rows
