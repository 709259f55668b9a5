"""Weight re-packing: reference ``state_dict`` tensors (fp32, torch layouts) -> bf16 K-major tiles that dn_gemm's
TMA/tcgen05 pipeline consumes.  Runs once at model load (host-side torch indexing only; no arithmetic beyond the
fp32 sum of the skip biases).

Packed layout rules (see include/diffnorm_b200.h, dn_gemm_desc):
  * K (input channel) extents are padded to multiples of 64 with zero columns; N to multiples of 16.
  * one N tile = 256 packed rows; plain GEMMs simply use consecutive rows;
    GEGLU tiles  = 128 "x" rows followed by the 128 matching "gate" rows (LM:881-885 chunk order);
    WaveNet tiles = 128 dilated-conv rows followed by the 128 matching res_conv rows (LM:509-510).
  * a K=3 causal conv is three K segments (taps) over the same activation tile shifted by 2d, d, 0 frames
    (LM:476-488: tap k multiplies x[t - (2-k) d]).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib
from .ops import GemmPlan

BK = 64
WT = 256


def rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _bf(t: torch.Tensor, fmt: str = "bf16") -> torch.Tensor:
    """fp32 packed weight -> the plan's operand format: bf16, fp16 (saturating), or the split pair [hi | lo] along K."""
    if fmt == "f16":
        return t.clamp(-65504.0, 65504.0).to(torch.float16).contiguous()
    hi = t.to(torch.bfloat16)      # "bf16" and "bf16sr" (the latter also keeps the fp32 master, see GemmPlan.W32)
    if fmt == "split":
        return torch.cat([hi, (t - hi.float()).to(torch.bfloat16)], dim=1).contiguous()
    return hi.contiguous()


def _mk(Wp: torch.Tensor, wfmt: str, *args, **kw) -> GemmPlan:
    """GemmPlan over the packed fp32 matrix Wp cast to the plan's operand format; "bf16sr" also keeps Wp itself (the master
    the engine re-rounds stochastically every sampler step)."""
    plan = GemmPlan(_bf(Wp, wfmt), *args, **kw)
    if wfmt == "bf16sr":
        plan.W32 = Wp.contiguous()
    return plan


def pack_linear(W: torch.Tensor, bias: Optional[torch.Tensor], epi: int = _lib.EPI_BF16, k_pad: Optional[int] = None,
                n_pad: Optional[int] = None, name: str = "", fmt: str = "bf16") -> GemmPlan:
    """W [N, K] (nn.Linear / 1x1 conv squeezed).  Output columns = n_pad (extra columns compute to bias 0)."""
    W = W.reshape(W.shape[0], -1).float()
    N, K = W.shape
    Kp = rup(K, BK) if k_pad is None else k_pad
    Np = rup(N, 16) if n_pad is None else n_pad
    Wp = torch.zeros(Np, Kp, device=W.device)
    Wp[:N, :K] = W
    bp = None
    if bias is not None:
        bp = torch.zeros(Np, device=W.device)
        bp[:N] = bias.float()
    return _mk(Wp, fmt, [(0, 0, Kp // BK, 0, 0)], Np, (Np + WT - 1) // WT, epi, bias=bp, name=name, fmt=fmt)


def pack_conv3(W: torch.Tensor, bias: Optional[torch.Tensor], epi: int = _lib.EPI_BF16, cin_pad: Optional[int] = None,
               n_pad: Optional[int] = None, dilation: int = 1, name: str = "", shift_sign: int = 1,
               fmt: str = "bf16") -> GemmPlan:
    """W [N, Cin, 3] causal conv -> K = [tap0 | tap1 | tap2], shifts (2d, d, 0).
    shift_sign = -1 reads x[t + (2-k) d] instead: the data-gradient of the conv when W is the (Cin, Cout)-transposed
    kernel (dx[t] = sum_k W_k^T dy[t + (2-k) d]; frames past the utterance read as zero)."""
    N, Cin, Kk = W.shape
    assert Kk == 3
    Cp = rup(Cin, BK) if cin_pad is None else cin_pad
    Np = rup(N, 16) if n_pad is None else n_pad
    Wp = torch.zeros(Np, 3 * Cp, device=W.device)
    for k in range(3):
        Wp[:N, k * Cp:k * Cp + Cin] = W[:, :, k].float()
    bp = None
    if bias is not None:
        bp = torch.zeros(Np, device=W.device)
        bp[:N] = bias.float()
    segs = [(0, shift_sign * (2 - k), Cp // BK, k * Cp, 0) for k in range(3)]
    return _mk(Wp, fmt, segs, Np, (Np + WT - 1) // WT, epi, bias=bp, dilation=dilation, name=name, fmt=fmt)


def pack_geglu(W: torch.Tensor, bias: torch.Tensor, name: str = "", fmt: str = "bf16") -> GemmPlan:
    """W [2*inner, K]: rows [0, inner) = x, [inner, 2 inner) = gate.  Output = inner_pad (multiple of 128) columns;
    padded lanes have zero weights and bias so they produce gelu(0) * 0 = 0 exactly."""
    W = W.float()
    inner, K = W.shape[0] // 2, W.shape[1]
    Kp = rup(K, BK)
    ip = rup(inner, 128)
    tiles = ip // 128
    Wp = torch.zeros(tiles * WT, Kp, device=W.device)
    bp = torch.zeros(tiles * WT, device=W.device)
    for j in range(tiles):
        lo, hi = j * 128, min((j + 1) * 128, inner)
        if hi <= lo:
            continue
        n = hi - lo
        Wp[j * WT:j * WT + n, :K] = W[lo:hi]
        Wp[j * WT + 128:j * WT + 128 + n, :K] = W[inner + lo:inner + hi]
        bp[j * WT:j * WT + n] = bias[lo:hi].float()
        bp[j * WT + 128:j * WT + 128 + n] = bias[inner + lo:inner + hi].float()
    return _mk(Wp, fmt, [(0, 0, Kp // BK, 0, 0)], ip, tiles, _lib.EPI_GEGLU, bias=bp, name=name, fmt=fmt)


def pack_wavenet_level(convs: List[torch.Tensor], conv_b: List[torch.Tensor], ress: List[torch.Tensor],
                       res_b: List[torch.Tensor], c_pad: int, name: str = "", fmt: str = "bf16") -> GemmPlan:
    """One WaveNet stack level = `len(convs)` independent chains (group g has dilation 2^g, LM:553-566), fused
    conv(k3, dilated) + res_conv(1x1) + FiLM/gate epilogue.  convs[g] [C, C, 3], ress[g] [C, C, 1].
    K layout per chain: [tap2 (shift 0, also feeds the res rows) | tap0 (shift 2d) | tap1 (shift d)]."""
    G = len(convs)
    Cc = convs[0].shape[0]
    Cp = c_pad
    assert Cp % 128 == 0 and Cp >= Cc
    tiles = Cp // 128
    rows_g = tiles * WT
    dev = convs[0].device
    Wp = torch.zeros(G * rows_g, 3 * Cp, device=dev)
    bc = torch.zeros(G * Cp, device=dev)
    br = torch.zeros(G * Cp, device=dev)
    order = (2, 0, 1)  # K position -> conv tap
    if Cc == Cp:  # vectorised form (denoiser: C = 512); same layout as the loop below
        cw = torch.stack([c.float() for c in convs])                       # [G, C, C, 3]
        rw = torch.stack([r.float().reshape(Cc, Cc) for r in ress])        # [G, C, C]
        v = Wp.view(G, tiles, 2, 128, 3 * Cp)
        for pos, tap in enumerate(order):
            v[:, :, 0, :, pos * Cp:(pos + 1) * Cp] = cw[..., tap].reshape(G, tiles, 128, Cc)
        v[:, :, 1, :, 0:Cp] = rw.reshape(G, tiles, 128, Cc)
        bc = torch.cat([b.float() for b in conv_b])
        br = torch.cat([b.float() for b in res_b])
    else:
        for g in range(G):
            cw, rw = convs[g].float(), ress[g].float().reshape(Cc, Cc)
            for j in range(tiles):
                lo, hi = j * 128, min((j + 1) * 128, Cc)
                if hi <= lo:
                    continue
                n = hi - lo
                r0 = g * rows_g + j * WT
                for pos, tap in enumerate(order):
                    Wp[r0:r0 + n, pos * Cp:pos * Cp + Cc] = cw[lo:hi, :, tap]
                Wp[r0 + 128:r0 + 128 + n, 0:Cc] = rw[lo:hi]
            bc[g * Cp:g * Cp + Cc] = conv_b[g].float()
            br[g * Cp:g * Cp + Cc] = res_b[g].float()
    kb = Cp // BK
    segs = [(0, 0, kb, 0, 0), (0, 2, kb, Cp, 128), (0, 1, kb, 2 * Cp, 128)]
    return _mk(Wp, fmt, segs, Cp, tiles, _lib.EPI_WN_GATE, bias=bc, bias2=br, groups=G, g_w_row=rows_g,
                    g_bias=Cp, dilation=1, dilation_shl_group=1, name=name, fmt=fmt)


def pack_skip_sum(skips: List[torch.Tensor], skip_b: List[torch.Tensor], c_pad: int, name: str = "",
                  fmt: str = "bf16") -> GemmPlan:
    """sum_g skip_conv_g(y_g) (LM:580,617) = one GEMM over the concatenated chain outputs [.., G * c_pad].
    Writes all c_pad output columns (the pad columns get exact zeros)."""
    G = len(skips)
    Cc = skips[0].shape[0]
    Wp = torch.zeros(c_pad, G * c_pad, device=skips[0].device)
    for g in range(G):
        Wp[:Cc, g * c_pad:g * c_pad + Cc] = skips[g].float().reshape(Cc, Cc)
    bp = torch.zeros(c_pad, device=skips[0].device)
    bp[:Cc] = torch.stack([b.float() for b in skip_b]).sum(0)
    Np = Wp.shape[0]
    return _mk(Wp, fmt, [(0, 0, G * c_pad // BK, 0, 0)], Np, (Np + WT - 1) // WT, _lib.EPI_BF16, bias=bp, name=name,
                    fmt=fmt)


# ------------------------------------------------------------------------------------------------ training-step packings
def geglu_row_map(inner: int) -> torch.Tensor:
    """Packed GEGLU row r (tile j: 128 x rows, 128 gate rows) -> row of the reference Linear weight [2*inner, K], or -1
    for a padding row."""
    ip = rup(inner, 128)
    m = torch.full((2 * ip,), -1, dtype=torch.long)
    for j in range(ip // 128):
        lo, hi = j * 128, min((j + 1) * 128, inner)
        if hi > lo:
            m[j * WT:j * WT + hi - lo] = torch.arange(lo, hi)
            m[j * WT + 128:j * WT + 128 + hi - lo] = inner + torch.arange(lo, hi)
    return m


def pack_wavenet_level_dgrad(convs: List[torch.Tensor], ress: List[torch.Tensor], c_pad: int, name: str = "") -> GemmPlan:
    """Data-gradient of one WaveNet level for all chains: input dur [.., G*2C] per chain [du | dres], output da_g =
    sum_k conv_k^T du[t + (2-k) d_g] + res^T dres[t].  K layout per chain: [conv_2^T | res^T | conv_0^T | conv_1^T]."""
    G = len(convs)
    Cc, Cp = convs[0].shape[0], c_pad
    dev = convs[0].device
    Wp = torch.zeros(G, Cp, 4 * Cp, device=dev)
    cw = torch.stack([c.float() for c in convs])                       # [G, Cout, Cin, 3]
    rw = torch.stack([r.float().reshape(Cc, Cc) for r in ress])
    Wp[:, :Cc, 0:Cc] = cw[..., 2].transpose(1, 2)
    Wp[:, :Cc, Cp:Cp + Cc] = rw.transpose(1, 2)
    Wp[:, :Cc, 2 * Cp:2 * Cp + Cc] = cw[..., 0].transpose(1, 2)
    Wp[:, :Cc, 3 * Cp:3 * Cp + Cc] = cw[..., 1].transpose(1, 2)
    Wp = Wp.view(G * Cp, 4 * Cp)
    kb = Cp // BK
    segs = [(0, 0, 2 * kb, 0, 0), (0, -2, kb, 2 * Cp, 0), (0, -1, kb, 3 * Cp, 0)]
    return GemmPlan(_bf(Wp), segs, Cp, (Cp + WT - 1) // WT, _lib.EPI_BF16, groups=G, g_w_row=Cp, g_bias=0, dilation=1,
                    dilation_shl_group=1, name=name)
