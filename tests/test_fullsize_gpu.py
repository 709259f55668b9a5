"""BASELINE config-2 size (64 utterances x 1000 frames) where the CPU oracle cannot follow in seconds: the pass is
checked through size-independent properties — determinism, independence of an utterance's result from the batch it
rides in (padding / batching invariance: no cross-utterance op exists on the path), agreement of the
tensor-core GEMM path with the one-thread-per-output SIMT checker kernel on one full-size denoiser call, and the
run-length invariants of `_reduce_tgt` on the full-size unit matrix."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from diffnorm_b200 import _lib  # noqa: E402
from diffnorm_b200.engine import DiffNormEngine  # noqa: E402
from oracle import diffnorm_oracle as O  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def big():
    arch = O.Arch(latent_dim=16)
    sd = O.init_state_dict(arch, seed=1, gains=O.PARITY_GAINS)
    eng = DiffNormEngine(sd, DEV)
    g = torch.Generator().manual_seed(99)
    B, T = 64, 1000
    lens = torch.randint(200, T + 1, (B,), generator=g)
    lens[0] = T
    mask = O.lengths_to_mask(lens, T)
    feat = (torch.randn(B, T, 768, generator=g) * mask[:, :, None]).to(DEV)
    ev, eq = torch.randn(B, 16, T, generator=g).to(DEV), torch.randn(B, T, 16, generator=g).to(DEV)
    return eng, feat, lens, ev, eq


def test_fullsize_determinism_batch_invariance_and_reduce_invariants(big):
    eng, feat, lens, ev, eq = big
    B, T, _ = feat.shape
    ld = lens.to(torch.int32).to(DEV)
    start = 6
    a = eng.normalize(feat, ld, start, ev, eq)
    x0_a, units_a = a["x0"].clone(), a["units"].clone()
    dedup, dur, keep, counts = (a[k].clone() for k in ("dedup", "duration", "index_to_keep", "counts"))
    b = eng.normalize(feat, ld, start, ev, eq)
    assert torch.equal(x0_a, b["x0"]) and torch.equal(units_a, b["units"])            # run-to-run determinism
    # the same 6 utterances alone, padded only to their own maximum length
    idx = [3, 17, 0, 42, 63, 8]
    Ts = int(lens[idx].max())
    sub = eng.normalize(feat[idx, :Ts].contiguous(), ld[idx].contiguous(), start, ev[idx, :, :Ts].contiguous(),
                        eq[idx, :Ts].contiguous())
    worst, same, tot = 0.0, 0, 0
    for j, i in enumerate(idx):
        n = int(lens[i])
        worst = max(worst, float((sub["x0"][j, :n] - x0_a[i, :n]).abs().max()))
        same += int((sub["units"][j, :n] == units_a[i, :n]).sum())
        tot += n
    print(f"[parity] full-size batch invariance: max |dx0| {worst:.3e}, units equal {same}/{tot}")
    assert worst <= 1e-5 and same >= 0.999 * tot
    # run-length invariants (repr_to_repr_unit_dataset.py:92-113) on the 64 x 1000 unit matrix
    u, dd, du, kk, cc = units_a.cpu(), dedup.cpu(), dur.cpu(), keep.cpu(), counts.cpu()
    for i in range(B):
        n, r = int(lens[i]), int(cc[i])
        assert int(du[i, :r].sum()) == n
        assert torch.equal(u[i, kk[i, :r]], dd[i, :r])
        assert r <= 1 or bool((dd[i, 1:r] != dd[i, : r - 1]).all())
        assert torch.equal(torch.repeat_interleave(dd[i, :r], du[i, :r]), u[i, :n])


def test_fullsize_denoiser_call_tcgen05_vs_simt_checker(big):
    eng, feat, lens, ev, eq = big
    B, T, _ = feat.shape
    ld = lens.to(torch.int32).to(DEV)
    x = torch.randn(B, T, 16, generator=torch.Generator().manual_seed(5)).to(DEV)
    xb = eng.stage_latent(x)
    t_idx = torch.tensor([57], dtype=torch.int32, device=DEV)
    got = eng.denoise(xb, ld, B, T, t_idx).view(B, T, -1)[..., :16].clone()
    eng.gemm_impl = _lib.GEMM_SIMT_CHECK
    try:
        want = eng.denoise(xb, ld, B, T, t_idx).view(B, T, -1)[..., :16].clone()
    finally:
        eng.gemm_impl = None
    mask = O.lengths_to_mask(lens, T).to(DEV)
    d = (got - want)[mask].abs()
    print(f"[parity] full-size eps_hat tcgen05 vs SIMT GEMMs: max {float(d.max()):.3e} mean {float(d.mean()):.3e} "
          f"ref std {float(want[mask].std()):.3e}")
    assert float(d.max()) <= 3e-2 * float(want[mask].std()) + 1e-3
