"""Test infrastructure: the fp32 oracle (oracle/diffnorm_oracle.py, pinned to the live reference) run on the GPU so that
full-size cases (BASELINE configs C1 / C2: 99 denoiser calls over thousands of frames) finish in seconds instead of the
minutes the host cores need.  Strict fp32: TF32 is switched off for matmul and cuDNN while the oracle runs, so this is
the same arithmetic the CPU oracle does (tests/test_fullsize_parity_gpu.py checks one small case of this against the
CPU oracle).  Never imported by the product."""
from __future__ import annotations

import contextlib
from typing import Dict, Optional

import torch

from oracle import diffnorm_oracle as O


@contextlib.contextmanager
def strict_fp32():
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    p = torch.get_float32_matmul_precision()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    orig_pe = O.pos_embed
    # the oracle builds its sinusoid table on the host; keep its arithmetic, place the result next to the mask
    O.pos_embed = lambda m, dim, dtype=torch.float32: orig_pe(m.cpu(), dim, dtype).to(m.device)
    try:
        yield
    finally:
        O.pos_embed = orig_pe
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b
        torch.set_float32_matmul_precision(p)


@torch.no_grad()
def oracle_pass(sd: Dict[str, torch.Tensor], arch: "O.Arch", feat, mask, start_step: int, eps_vae, eps_q,
                keep_steps=(), chunk: Optional[int] = None) -> Dict[str, object]:
    """O.normalize_pass (sampler 'ddim', LM:1386-1471) with every tensor on feat.device.  Also returns the latent
    entering the denoiser call at each t in `keep_steps` ("x_at") and the eps_hat of that call ("eps_at").
    chunk: utterances per oracle call (the reference materialises the [B, 8, N, N] attention matrix)."""
    dev = feat.device
    sch = O.Schedule(arch.timesteps)
    B = feat.shape[0]
    chunk = chunk or B
    out: Dict[str, object] = {"x_at": {}, "eps_at": {}}
    with strict_fp32():
        z = torch.cat([O.vae_encode(sd, arch, feat[i:i + chunk], eps_vae[i:i + chunk]) for i in range(0, B, chunk)])
        x = O.q_sample(sch, z, start_step, eps_q)
        out["z"], out["x_start"] = z, x.clone()
        for t in (range(start_step - 1, 0, -1) if start_step > 1 else [0]):   # LM:1402,1444
            eh = torch.cat([O.denoiser(sd, arch, x[i:i + chunk], torch.full((min(chunk, B - i),), t, dtype=torch.long,
                                                                            device=dev), mask[i:i + chunk])
                            for i in range(0, B, chunk)])
            if t in keep_steps:
                out["x_at"][t], out["eps_at"][t] = x.clone(), eh.clone()
            x = O.ddim_step(sch, x, eh, t)
        out["x0"] = x
        dec = [O.vae_decode(sd, arch, x[i:i + chunk], mask[i:i + chunk]) for i in range(0, B, chunk)]
        out["recon"] = torch.cat([d[0] for d in dec])
        out["logits"] = torch.cat([d[1] for d in dec])
    out["units"] = torch.argmax(out["logits"], dim=-1) - O.UNIT_OFFSET
    return out


@torch.no_grad()
def oracle_decode(sd, arch, latent, mask, chunk: Optional[int] = None):
    B = latent.shape[0]
    chunk = chunk or B
    with strict_fp32():
        dec = [O.vae_decode(sd, arch, latent[i:i + chunk], mask[i:i + chunk]) for i in range(0, B, chunk)]
    return torch.cat([d[0] for d in dec]), torch.cat([d[1] for d in dec])


@torch.no_grad()
def oracle_denoise(sd, arch, x, t: int, mask, chunk: Optional[int] = None):
    B = x.shape[0]
    chunk = chunk or B
    with strict_fp32():
        return torch.cat([O.denoiser(sd, arch, x[i:i + chunk], torch.full((min(chunk, B - i),), t, dtype=torch.long,
                                                                          device=x.device), mask[i:i + chunk])
                          for i in range(0, B, chunk)])


def ragged_lengths(B: int, T: int, seed: int = 3):
    """One full-length utterance, the rest uniform in [T/2, T] (the driver pads a batch to its longest member)."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    lens[0] = T
    return lens


def case_inputs(z: int, B: int, T: int, seed: int = 7, ragged: bool = True):
    g = torch.Generator().manual_seed(seed)
    lens = ragged_lengths(B, T) if ragged else torch.full((B,), T)
    mask = O.lengths_to_mask(lens, T)
    feat = torch.randn(B, T, 768, generator=g) * mask[:, :, None]
    return dict(lens=lens, mask=mask, feat=feat, eps_vae=torch.randn(B, z, T, generator=g),
                eps_q=torch.randn(B, T, z, generator=g))
