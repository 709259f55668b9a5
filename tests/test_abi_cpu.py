"""CPU-side checks of the boundary: the C-ABI library loads and exports exactly what include/diffnorm_b200.h
declares (no compute calls without a GPU), argument validation returns DN_EINVAL, host-side packing and schedule
logic agree with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    so = os.path.join(ROOT, "diffnorm_b200", "csrc", "libdiffnorm_b200.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    return so


def test_library_exports_every_declared_symbol():
    so = _ensure_built()
    hdr = open(os.path.join(ROOT, "include", "diffnorm_b200.h")).read()
    declared = sorted(set(re.findall(r"^\s*(?:int|int64_t|unsigned long long)\s+(dn_[a-z0-9_]+)\s*\(", hdr, flags=re.M)))
    assert len(declared) >= 18
    lib = ctypes.CDLL(so)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    from diffnorm_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared  # the ctypes binding covers the whole header
    assert _lib.lib.dn_abi_version() == _lib.ABI_VERSION == 2


def test_argument_errors_are_reported_not_crashed():
    _ensure_built()
    from diffnorm_b200 import _lib
    L = _lib.lib
    assert L.dn_reduce_tgt(None, None, 1, 1, None, None, None, None, None) == -1
    assert L.dn_argmax_units(None, 0, 1, 4, 4, 4, None, None) == -1
    assert L.dn_gemm(None, 0, None) == -1
    assert L.dn_attention(None, None, None, 1, 1, 1, 64, 0, 0, None) == -1
    d = _lib.GemmDesc()
    assert L.dn_gemm(ctypes.byref(d), 0, None) == -1
    with pytest.raises(_lib.DiffNormLibraryError):
        _lib.check(-1, "x")


def test_ops_reject_cpu_tensors():
    _ensure_built()
    from diffnorm_b200 import ops
    with pytest.raises(ValueError, match="CUDA"):
        ops.reduce_tgt(torch.zeros(1, 4, dtype=torch.int64), torch.ones(1, dtype=torch.int32))


def test_schedule_rows_match_oracle():
    from diffnorm_b200.schedule import DDPMScheduler
    from oracle import diffnorm_oracle as O
    s, o = DDPMScheduler(200), O.Schedule(200)
    for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "posterior_mean_coef1", "posterior_log_variance_clipped"):
        np.testing.assert_array_equal(getattr(s, k), getattr(o, k))
    rows = s.ddim_rows()
    x, e = torch.randn(5, 7), torch.randn(5, 7)
    for t in (150, 99, 1):
        c = torch.from_numpy(rows[t])
        x0 = (x - c[1] * e) / c[0].clamp(min=1e-10)
        pn = (x - c[0] * x0) / c[1].clamp(min=1e-10)
        torch.testing.assert_close(x0 * c[2] + c[3] * pn, O.ddim_step(o, x, e, t), rtol=1e-6, atol=1e-6)
    sp, tmap = s.spaced(range(0, 40, 4))
    so, tmo = O.Schedule.spaced(o, range(0, 40, 4))
    assert tmap == tmo
    np.testing.assert_allclose(sp.betas, so.betas, rtol=1e-14)
    d = s.ddpm_rows()
    assert d[0, 4] == 0.0 and d[1, 4] > 0


def test_packing_layouts():
    _ensure_built()
    from diffnorm_b200 import packing
    g = torch.Generator().manual_seed(0)
    W = torch.randn(10, 20, 3, generator=g)
    p = packing.pack_conv3(W, torch.zeros(10))
    assert p.W.shape == (16, 3 * 64) and [s[1] for s in p.segs] == [2, 1, 0]
    assert torch.equal(p.W[:10, 64:84].float(), W[:, :, 1].bfloat16().float())
    Wg, bg = torch.randn(2 * 200, 64, generator=g), torch.randn(400, generator=g)
    pg = packing.pack_geglu(Wg, bg)
    assert pg.W.shape == (512, 64) and pg.n_out == 256 and pg.n_tiles == 2
    assert torch.equal(pg.W[256 + 128:256 + 128 + 72].float(), Wg[200 + 128:].bfloat16().float())  # gate rows of tile 1
    assert (pg.W[256 + 72:256 + 128] == 0).all() and (pg.bias[256 + 72:256 + 128] == 0).all()
    conv = [torch.randn(192, 192, 3, generator=g) for _ in range(2)]
    res = [torch.randn(192, 192, 1, generator=g) for _ in range(2)]
    b = [torch.randn(192, generator=g) for _ in range(2)]
    pw = packing.pack_wavenet_level(conv, b, res, b, 256)
    assert pw.W.shape == (2 * 512, 768) and pw.groups == 2 and pw.g_w_row == 512
    # tile 1 of chain 1: conv rows 128..191 at K position 0 hold tap 2 (shift 0); res rows follow at +128
    r0 = 512 + 256
    assert torch.equal(pw.W[r0:r0 + 64, :192].float(), conv[1][128:, :, 2].bfloat16().float())
    assert torch.equal(pw.W[r0:r0 + 64, 256:448].float(), conv[1][128:, :, 0].bfloat16().float())
    assert torch.equal(pw.W[r0 + 128:r0 + 192, :192].float(), res[1][128:, :, 0].bfloat16().float())
    assert (pw.W[r0 + 128:r0 + 256, 256:] == 0).all()
    assert [tuple(s) for s in pw.segs] == [(0, 0, 4, 0, 0), (0, 2, 4, 256, 128), (0, 1, 4, 512, 128)]
    # operand formats: fp16 weights (the sampler loop) and split-precision pairs [hi | lo] along K (the VAE)
    ph = packing.pack_conv3(W, torch.zeros(10), fmt="f16")
    assert ph.W.dtype == torch.float16 and torch.equal(ph.W[:10, 64:84].float(), W[:, :, 1].half().float())
    ps = packing.pack_conv3(W, torch.zeros(10), fmt="split")
    assert ps.W.shape == (16, 2 * 3 * 64) and ps.W.dtype == torch.bfloat16 and ps.fmt == "split"
    hi, lo = ps.W[:10, 64:84].float(), ps.W[:10, 192 + 64:192 + 84].float()
    assert torch.equal(hi, W[:, :, 1].bfloat16().float())
    assert (hi + lo - W[:, :, 1]).abs().max() <= 2.0 ** -16 * W.abs().max()
