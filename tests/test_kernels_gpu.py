"""Per-kernel parity tests (B200 only): every C-ABI entry point against a plain torch fp32 statement of the same
op / the oracle, on seeded inputs.  Integer kernels are bit-exact; bf16 tensor-core kernels carry the tolerance
in the test."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from diffnorm_b200 import _lib, ops, packing  # noqa: E402
from diffnorm_b200.schedule import DDPMScheduler  # noqa: E402
from oracle import diffnorm_oracle as O  # noqa: E402

DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


# --------------------------------------------------------------------------------------------------- integer kernels
def test_reduce_tgt_golden_bit_exact():
    g = np.load(os.path.join(GOLD, "reduce_tgt.npz"))
    names = sorted({f[:-3] for f in g.files if f.endswith("_in")})
    T = max(max(len(g[n + "_in"]) for n in names), 1)
    units = torch.zeros(len(names), T, dtype=torch.int64)
    lens = torch.zeros(len(names), dtype=torch.int32)
    for i, n in enumerate(names):
        a = g[n + "_in"]
        units[i, : len(a)] = torch.from_numpy(a)
        lens[i] = len(a)
    d, du, kp, cnt = ops.reduce_tgt(units.to(DEV), lens.to(DEV))
    d, du, kp, cnt = d.cpu(), du.cpu(), kp.cpu(), cnt.cpu()
    for i, n in enumerate(names):
        r = int(cnt[i])
        assert d[i, :r].tolist() == g[n + "_dedup"].tolist(), n
        assert kp[i, :r].tolist() == g[n + "_keep"].tolist(), n
        nd = len(g[n + "_dur"])
        assert du[i, :nd].tolist() == g[n + "_dur"].tolist(), n


def test_reduce_tgt_large_properties():
    rng = np.random.default_rng(0)
    B, T = 64, 2000
    runs = rng.geometric(0.6, size=(B, T))
    ids = rng.integers(0, 1000, size=(B, T))
    units = np.stack([np.repeat(ids[b], runs[b])[:T] for b in range(B)])
    lens = rng.integers(1, T + 1, size=B).astype(np.int32)
    lens[0], lens[1] = T, 1
    d, du, kp, cnt = (t.cpu().numpy() for t in ops.reduce_tgt(torch.from_numpy(units).to(DEV), torch.from_numpy(lens).to(DEV)))
    for b in range(B):
        e_d, e_du, e_k = O.reduce_tgt_np(units[b, : lens[b]])
        r = cnt[b]
        assert r == len(e_d)
        assert (d[b, :r] == e_d).all() and (du[b, :r] == e_du).all() and (kp[b, :r] == e_k).all()
        assert du[b, :r].sum() == lens[b]                       # durations partition the utterance
        assert (np.repeat(d[b, :r], du[b, :r]) == units[b, : lens[b]]).all()  # expand(reduce(x)) == x
    # idempotence: reducing the reduced stream changes nothing
    d2, du2, kp2, cnt2 = ops.reduce_tgt(torch.from_numpy(d).to(DEV), torch.from_numpy(cnt.astype(np.int32)).to(DEV))
    assert (cnt2.cpu().numpy() == cnt).all()
    assert all((du2[b, : cnt[b]].cpu().numpy() == 1).all() for b in range(B))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_argmax_units(dtype):
    x = rnd(3, 50, 1008, seed=1).to(dtype)
    x[0, 0, :] = 0.0                 # all ties -> index 0
    x[0, 1, 7] = x[0, 1, 900] = 50   # tie -> first
    x[0, 2, 1003] = 60               # last valid class
    x[0, 3, 1005] = 99               # pad column must be ignored
    x[0, 4, 2] = 70                  # special symbol wins -> negative unit
    got = ops.argmax_units(x.contiguous(), 1004, 4).cpu()
    want = (torch.argmax(x[..., :1004].float(), dim=-1) - 4).cpu()
    assert torch.equal(got, want)
    assert got[0, 0] == -4 and got[0, 1] == 3 and got[0, 2] == 999 and got[0, 4] == -2


def test_unit_accuracy_and_gather():
    B, T = 5, 77
    g = torch.Generator().manual_seed(2)
    u = torch.randint(0, 5, (B, T), generator=g)
    r = torch.randint(0, 5, (B, T), generator=g)
    lens = torch.tensor([77, 1, 30, 64, 65], dtype=torch.int32)
    out = ops.unit_accuracy(u.to(DEV), r.to(DEV), lens.to(DEV)).cpu()
    m = torch.arange(T)[None] < lens[:, None]
    assert out.tolist() == [int(((u == r) & m).sum()), int(m.sum())]
    # gather + pad
    C = 768
    src = rnd(400, C, seed=3)
    row0 = torch.tensor([0, 100, 150, 200, 300], dtype=torch.int64)
    keep = torch.zeros(B, T, dtype=torch.int64)
    cnt = torch.tensor([10, 0, 5, 77, 3], dtype=torch.int32)
    for b in range(B):
        keep[b, : cnt[b]] = torch.sort(torch.randperm(90, generator=g)[: cnt[b]]).values
    for dt in (torch.float32, torch.bfloat16):
        dst = ops.gather_pack(src, row0.to(DEV), keep.to(DEV), cnt.to(DEV), T, dst_dtype=dt).float().cpu()
        for b in range(B):
            want = src.cpu()[row0[b] + keep[b, : cnt[b]]]
            if dt == torch.bfloat16:
                want = want.bfloat16().float()
            assert torch.equal(dst[b, : cnt[b]], want)
            assert (dst[b, cnt[b]:] == 0).all()


# --------------------------------------------------------------------------------------------------- elementwise
def test_diffusion_steps_match_oracle():
    sch_o, sch = O.Schedule(200), DDPMScheduler(200)
    B, T, z = 3, 50, 16
    x, e, n = rnd(B, T, z, seed=4), rnd(B, T, z, seed=5), rnd(B, T, z, seed=6)
    ddim_rows = torch.from_numpy(sch.ddim_rows()).to(DEV)
    t_idx = torch.zeros(1, dtype=torch.int32, device=DEV)
    for t in (150, 99, 1):
        t_idx.fill_(t)
        xx = x.clone()
        xb = torch.full((B * T, 64), 7.0, dtype=torch.bfloat16, device=DEV)
        ops.ddim_step(xx, e, ddim_rows, t_idx, 0, xb)
        want = O.ddim_step(sch_o, x.cpu(), e.cpu(), t)
        torch.testing.assert_close(xx.cpu(), want, rtol=2e-5, atol=2e-6)
        torch.testing.assert_close(xb[:, :z].float().cpu().view(B, T, z), want.bfloat16().float(), rtol=1e-2, atol=1e-2)
        xx = x.clone()
        ops.ddim_step(xx, e, ddim_rows, t_idx, 1, None)
        torch.testing.assert_close(xx.cpu(), O.ddim_generic_step(sch_o, x.cpu(), e.cpu(), t), rtol=2e-5, atol=2e-6)
        for large in (False, True):
            rows = torch.from_numpy(sch.ddpm_rows(large)).to(DEV)
            xx = x.clone()
            ops.ddpm_step(xx, e, n, rows, t_idx, None)
            torch.testing.assert_close(xx.cpu(), O.ddpm_step(sch_o, x.cpu(), e.cpu(), t, n.cpu(), large), rtol=2e-5, atol=2e-6)
    ops.advance_step(t_idx, -1)
    assert int(t_idx) == 0
    # q_sample + bf16 staging with zero pad
    zl, eq = rnd(B, T, z, seed=7), rnd(B, T, z, seed=8)
    xo = torch.empty(B * T, z, device=DEV)
    xb = torch.full((B * T, 64), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.q_sample(zl, eq, float(np.float32(sch.sqrt_alphas_cumprod[100])), float(np.float32(sch.sqrt_one_minus_alphas_cumprod[100])), xo, xb)
    torch.testing.assert_close(xo.view(B, T, z).cpu(), O.q_sample(sch_o, zl.cpu(), 100, eq.cpu()), rtol=1e-6, atol=1e-6)
    assert (xb[:, z:] == 0).all()


def test_vae_reparam_and_cast():
    B, T, z = 2, 33, 16
    params = rnd(B, T, 2 * z, seed=9, scale=3.0)
    params[0, 0, z] = 100.0   # clamp high
    params[0, 1, z] = -100.0  # clamp low
    eps_cf = rnd(B, z, T, seed=10)
    got = ops.vae_reparam(params, eps_cf, z, True).cpu()
    want = O.vae_sample(params.cpu().transpose(1, 2), eps_cf.cpu()).transpose(1, 2)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    got2 = ops.vae_reparam(params, eps_cf.transpose(1, 2).contiguous(), z, False).cpu()
    torch.testing.assert_close(got2, want, rtol=1e-5, atol=1e-5)
    src = rnd(70, 20, seed=11)
    out = ops.cast_pad_bf16(src, 64)
    assert torch.equal(out[:, :20].float(), src.bfloat16().float()) and (out[:, 20:] == 0).all()


@pytest.mark.parametrize("C", [512, 768])
def test_adarmsnorm(C):
    B, T = 3, 41
    x = rnd(B * T, C, seed=12, scale=2.0)
    x[5] = 0.0  # zero row: eps clamp
    gp = rnd(C, seed=13) + 1
    table = rnd(4, 3, 2 * C, seed=14)
    t_idx = torch.tensor([2, 0, 3], dtype=torch.int32, device=DEV)
    out = torch.empty(B * T, C, dtype=torch.bfloat16, device=DEV)
    ops.adarmsnorm(x, out, B, T, gp)
    want = O.rmsnorm(x.cpu().view(B, T, C), gp.cpu())
    torch.testing.assert_close(out.float().cpu().view(B, T, C), want, rtol=1e-2, atol=1e-2)
    gb = table.view(-1)[1 * 2 * C:]  # layer 1 of 3
    ops.adarmsnorm(x, out, B, T, None, gb, 3 * 2 * C, t_idx, 1)
    want = O.rmsnorm(x.cpu().view(B, T, C), None, table.cpu()[t_idx.cpu().long(), 1])
    torch.testing.assert_close(out.float().cpu().view(B, T, C), want, rtol=1e-2, atol=2e-2)


def test_wavenet_gate_kernel():
    B, T, C = 2, 30, 512
    u, r = rnd(B * T, C, seed=15, scale=2.0).bfloat16(), rnd(B * T, C, seed=16).bfloat16()
    table = rnd(5, 2 * C, seed=17)
    t_idx = torch.tensor([4], dtype=torch.int32, device=DEV)
    y = torch.empty_like(u)
    ops.wavenet_gate(u, r, y, B, T, table.view(-1), 2 * C, t_idx, 0)
    uu = u.float() * table[4, :C] + table[4, C:]
    want = uu.tanh() * uu.sigmoid() + r.float()
    torch.testing.assert_close(y.float(), want, rtol=1e-2, atol=1e-2)


def test_randn_kernel_is_standard_normal_and_reproducible():
    a = ops.randn((1000, 1003), DEV, 7, 0)
    b = ops.randn((1000, 1003), DEV, 7, 0)
    c = ops.randn((1000, 1003), DEV, 8, 0)
    d = ops.randn((1000, 1003), DEV, 7, 1)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, d)
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1.0) < 5e-3
    assert abs(float((a ** 3).mean())) < 2e-2 and abs(float((a ** 4).mean()) - 3.0) < 5e-2      # skewness 0, kurtosis 3
    assert abs(float((a[:, :-1] * a[:, 1:]).mean())) < 5e-3                                       # no neighbour correlation
    assert float(a.abs().max()) < 7.0 and bool(torch.isfinite(a).all())


def test_time_table_kernels():
    w = rnd(256, seed=18)
    steps = torch.arange(200, dtype=torch.int32, device=DEV)
    f = ops.time_features(steps, w).cpu()
    tf = torch.arange(200, dtype=torch.float)[:, None]
    fr = tf * w.cpu()[None] * 2 * np.pi
    want = torch.cat([tf, fr.sin(), fr.cos()], -1)
    torch.testing.assert_close(f, want, rtol=1e-4, atol=2e-3)  # sin/cos of arguments up to ~4e3 rad in fp32
    W, b, x = rnd(300, 513, seed=19, scale=0.05), rnd(300, seed=20), rnd(37, 513, seed=21)
    got = ops.linear_f32(x, W, b, act=1).cpu()
    want = torch.nn.functional.silu(x.cpu() @ W.cpu().T + b.cpu())
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


# --------------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("dh,T,lens", [(64, 200, [200, 131]), (96, 77, [77, 1]), (64, 1000, [1000, 640])])
def test_attention(dh, T, lens):
    B, H = 2, 8
    qkv = rnd(B, T, 3 * H * dh, seed=22).bfloat16()
    lengths = torch.tensor(lens, dtype=torch.int32, device=DEV)
    out = torch.empty(B, T, H * dh, dtype=torch.bfloat16, device=DEV)
    ops.attention(qkv, out, lengths, B, T, H, dh)
    q, k, v = (t.float().view(B, T, H, dh).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * dh ** -0.5
    mask = torch.arange(T, device=DEV)[None] < lengths[:, None]
    sim = sim.masked_fill(~mask[:, None, None, :], -torch.finfo(torch.float32).max)
    want = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), v).transpose(1, 2).reshape(B, T, H * dh)
    torch.testing.assert_close(out.float(), want, rtol=2e-2, atol=2e-2)


def test_attention_tcgen05_matches_mma_kernel():
    """dh = 64 runs on the tcgen05/TMEM kernel; DN_ATTN_IMPL=mma selects the mma.sync kernel, DN_ATTN_PERSIST=1 the persistent
    form of the tcgen05 kernel (attention_tcp.cu; lengths include an empty utterance and a single key): all must agree."""
    B, T, H, dh = 4, 700, 8, 64
    qkv = rnd(B, T, 3 * H * dh, seed=23, scale=1.5).bfloat16()
    lengths = torch.tensor([700, 129, 1, 0], dtype=torch.int32, device=DEV)
    outs = []
    for impl, persist in (("mma", "0"), ("tc", "0"), ("tc", "1")):
        os.environ["DN_ATTN_IMPL"], os.environ["DN_ATTN_PERSIST"] = impl, persist
        out = torch.full((B, T, H * dh), 9.0, dtype=torch.bfloat16, device=DEV)
        ops.attention(qkv, out, lengths, B, T, H, dh)
        torch.cuda.synchronize()
        outs.append(out.float())
    os.environ.pop("DN_ATTN_IMPL"), os.environ.pop("DN_ATTN_PERSIST")
    torch.testing.assert_close(outs[1], outs[0], rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(outs[2], outs[1], rtol=1e-2, atol=1e-2)     # same arithmetic up to the lazy-rescale decisions
    assert (outs[2][3] == 0).all() and (outs[1][3] == 0).all()


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_attention_optimistic_softmax_redo_path(dt):
    """Key blocks after the first exponentiate against the running max before the block's own max is known
    (attention_tc.cu OPT); a block whose max exceeds it by more than 2^8 is redone.  Scores that climb steeply from key
    block to key block force that path in every block; scores that fall never take it."""
    B, T, H, dh = 2, 640, 8, 64
    g = torch.Generator().manual_seed(5)
    for direction in (+1.0, -1.0):
        q = torch.randn(B, T, H * dh, generator=g) * 0.3 + 1.0
        ramp = (torch.arange(T) // 128).float()[None, :, None] * direction          # grows by one unit per key block
        k = torch.randn(B, T, H * dh, generator=g) * 0.3 + 0.9 * ramp              # s*scale*log2(e) moves ~10 per key block
        v = torch.randn(B, T, H * dh, generator=g)
        qkv = torch.cat([q, k, v], -1).to(DEV).to(dt).contiguous()
        lengths = torch.tensor([640, 389], dtype=torch.int32, device=DEV)
        out = torch.empty(B, T, H * dh, dtype=dt, device=DEV)
        ops.attention(qkv, out, lengths, B, T, H, dh)
        qq, kk, vv = (t.float().view(B, T, H, dh).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        sim = torch.einsum("bhid,bhjd->bhij", qq, kk) * dh ** -0.5
        mask = torch.arange(T, device=DEV)[None] < lengths[:, None]
        sim = sim.masked_fill(~mask[:, None, None, :], -torch.finfo(torch.float32).max)
        jump = (sim[0, 0, 0, 128:256].max() - sim[0, 0, 0, :128].max()) * 1.4427
        assert direction * float(jump) > 8.0                                       # the redo threshold really is crossed
        want = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), vv).transpose(1, 2).reshape(B, T, H * dh)
        torch.testing.assert_close(out.float(), want, rtol=2e-2, atol=2e-2)


def test_attention_dh96_tcgen05_matches_mma_kernel():
    """dh = 96 (VAE decoder) runs on the tcgen05/TMEM kernel of attention_tc96.cu (two 64-column boxes per operand);
    DN_ATTN_IMPL=mma selects the mma.sync kernel: forward (bf16) and the training form (dropout bits + row statistic)."""
    B, T, H, dh = 3, 700, 8, 96
    qkv = rnd(B, T, 3 * H * dh, seed=24, scale=1.2).bfloat16()
    lengths = torch.tensor([700, 129, 1], dtype=torch.int32, device=DEV)
    keep = torch.empty(B, H, T, (T + 31) // 32, dtype=torch.int32, device=DEV)
    ops.dropout_bits(keep, 0.1, 7, 0)
    res = {}
    for impl in ("mma", "tc"):
        os.environ["DN_ATTN_IMPL"] = impl
        out = torch.full((B, T, H * dh), 9.0, dtype=torch.bfloat16, device=DEV)
        ops.attention(qkv, out, lengths, B, T, H, dh)
        out_t = torch.full((B, T, H * dh), 9.0, dtype=torch.bfloat16, device=DEV)
        lse = torch.zeros(B * H, T, device=DEV)
        ops.attention_train(qkv, out_t, lse, lengths, keep, 1.0 / 0.9, B, T, H, dh)
        torch.cuda.synchronize()
        res[impl] = (out.float(), out_t.float(), lse.clone())
    os.environ.pop("DN_ATTN_IMPL")
    mask = (torch.arange(T, device=DEV)[None] < lengths[:, None])
    torch.testing.assert_close(res["tc"][0], res["mma"][0], rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(res["tc"][1], res["mma"][1], rtol=2e-2, atol=3e-2)
    lm = mask[:, None, :].expand(B, H, T).reshape(B * H, T)
    torch.testing.assert_close(res["tc"][2][lm], res["mma"][2][lm], rtol=1e-3, atol=2e-3)
    # and against the plain statement of the op
    q, k, v = (t.float().view(B, T, H, dh).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * dh ** -0.5
    sim = sim.masked_fill(~mask[:, None, None, :], -torch.finfo(torch.float32).max)
    want = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), v).transpose(1, 2).reshape(B, T, H * dh)
    torch.testing.assert_close(res["tc"][0], want, rtol=2e-2, atol=2e-2)


# --------------------------------------------------------------------------------------------------- GEMM
def _gemm_case(plan, A, out_shape, out_dtype, B, T, **kw):
    outs = []
    for impl in (_lib.GEMM_SIMT_CHECK, _lib.GEMM_TCGEN05):
        torch.manual_seed(0)
        out = torch.randn(out_shape, device=DEV).to(out_dtype).contiguous()  # RESID adds in place: same start
        plan.run(A, out, B, T, impl=impl, **kw)
        outs.append(out.float())
    torch.cuda.synchronize()
    # the CTA-pair (cta_group::2) form of the same kernel must reproduce the single-CTA result bit for bit
    torch.manual_seed(0)
    out2 = torch.randn(out_shape, device=DEV).to(out_dtype).contiguous()
    plan.run(A, out2, B, T, impl=_lib.GEMM_TCGEN05_2CTA, **kw)
    torch.cuda.synchronize()
    assert torch.equal(out2.float(), outs[1]), "cta_group::2 kernel differs from the single-CTA kernel"
    # packed rows (M tiles filled with row chunks across utterance boundaries) are only a different tiling: bit-identical
    for rc in (8, 16, 32, 64):
        for impl in (_lib.GEMM_TCGEN05, _lib.GEMM_TCGEN05_2CTA):
            torch.manual_seed(0)
            out3 = torch.randn(out_shape, device=DEV).to(out_dtype).contiguous()
            plan.run(A, out3, B, T, impl=impl, row_chunk=rc, **kw)
            assert torch.equal(out3.float(), outs[1]), f"row_chunk {rc} impl {impl} differs from the per-utterance tiling"
    return outs


def test_gemm_linear_epilogues():
    B, T, K, N = 2, 200, 192, 272
    A = rnd(B * T, K, seed=30).bfloat16()
    W, b = rnd(N, K, seed=31, scale=0.1), rnd(N, seed=32)
    want = (A.float() @ W.bfloat16().float().T + b).view(B * T, N)
    for epi, dt in ((_lib.EPI_BF16, torch.bfloat16), (_lib.EPI_F32, torch.float32), (_lib.EPI_RESID, torch.float32)):
        plan = packing.pack_linear(W.cpu(), b.cpu(), epi=epi).to(DEV)
        chk, tc = _gemm_case(plan, A, (B * T, N), dt, B, T)
        ref = want
        if epi == _lib.EPI_RESID:
            torch.manual_seed(0)
            ref = want + torch.randn((B * T, N), device=DEV)
        torch.testing.assert_close(chk, ref, rtol=1e-2, atol=2e-2)
        torch.testing.assert_close(tc, ref, rtol=1e-2, atol=2e-2)
        torch.testing.assert_close(tc, chk, rtol=1e-2, atol=1e-2)


def test_gemm_positional_epilogue():
    B, T, K, N = 2, 150, 64, 512
    A = rnd(B * T, K, seed=33).bfloat16()
    W, b = rnd(N, K, seed=34, scale=0.1), rnd(N, seed=35)
    lengths = torch.tensor([150, 90], dtype=torch.int32, device=DEV)
    pe = O.sinusoid_table(T + 1, N).to(DEV).contiguous()
    plan = packing.pack_linear(W.cpu(), b.cpu(), epi=_lib.EPI_F32).to(DEV)
    chk, tc = _gemm_case(plan, A, (B * T, N), torch.float32, B, T, pe=pe, lengths=lengths)
    mask = torch.arange(T, device=DEV)[None] < lengths[:, None]
    want = (A.float() @ W.bfloat16().float().T + b).view(B, T, N) + O.pos_embed(mask.cpu(), N).to(DEV)
    torch.testing.assert_close(tc.view(B, T, N), want, rtol=1e-2, atol=2e-2)
    torch.testing.assert_close(tc, chk, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("dil", [1, 4, 128])
def test_gemm_causal_conv(dil):
    B, T, Cin, N = 2, 300, 128, 144
    x = rnd(B, T, Cin, seed=36).bfloat16()
    W, b = rnd(N, Cin, 3, seed=37, scale=0.1), rnd(N, seed=38)
    plan = packing.pack_conv3(W.cpu(), b.cpu(), dilation=dil).to(DEV)
    chk, tc = _gemm_case(plan, x.view(B * T, Cin), (B * T, N), torch.bfloat16, B, T)
    want = O.causal_conv1d(x.float().transpose(1, 2), W.bfloat16().float(), b, dil).transpose(1, 2).reshape(B * T, N)
    torch.testing.assert_close(tc, want, rtol=2e-2, atol=3e-2)
    torch.testing.assert_close(tc, chk, rtol=1e-2, atol=1e-2)


def test_gemm_geglu():
    B, T, K, inner = 2, 130, 128, 200
    A = rnd(B * T, K, seed=39).bfloat16()
    W, b = rnd(2 * inner, K, seed=40, scale=0.1), rnd(2 * inner, seed=41)
    plan = packing.pack_geglu(W.cpu(), b.cpu()).to(DEV)
    assert plan.n_out == 256
    chk, tc = _gemm_case(plan, A, (B * T, 256), torch.bfloat16, B, T)
    h = A.float() @ W.bfloat16().float().T + b
    want = torch.nn.functional.gelu(h[:, inner:]) * h[:, :inner]
    torch.testing.assert_close(tc[:, :inner], want, rtol=2e-2, atol=2e-2)
    assert (tc[:, inner:] == 0).all()
    torch.testing.assert_close(tc, chk, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("cond", [False, True])
def test_gemm_wavenet_level(cond):
    B, T, C, G = 2, 260, 192, 3
    Cp = 256
    x = torch.zeros(B, T, Cp, device=DEV)
    x[..., :C] = rnd(B, T, C, seed=42)
    x = x.bfloat16()
    cw = [rnd(C, C, 3, seed=43 + g, scale=0.08) for g in range(G)]
    cb = [rnd(C, seed=50 + g) for g in range(G)]
    rw = [rnd(C, C, 1, seed=60 + g, scale=0.08) for g in range(G)]
    rb = [rnd(C, seed=70 + g) for g in range(G)]
    plan = packing.pack_wavenet_level([w.cpu() for w in cw], [w.cpu() for w in cb], [w.cpu() for w in rw],
                                      [w.cpu() for w in rb], Cp).to(DEV)
    kw = {}
    if cond:  # only meaningful when C == Cp; emulate with a table whose pad lanes are (gamma=1, beta=0)
        table = torch.zeros(4, G, 2 * Cp, device=DEV)
        table[..., :Cp] = 1.0
        table[..., :C] = rnd(4, G, C, seed=80) + 1
        table[..., Cp:Cp + C] = rnd(4, G, C, seed=81)
        t_idx = torch.tensor([3, 1], dtype=torch.int32, device=DEV)
        kw = dict(gb=table.view(-1), gb_t_stride=G * 2 * Cp, g_gb=2 * Cp, gb_half=Cp, t_idx=t_idx, t_idx_stride=1)
    chk, tc = _gemm_case(plan, x.view(B * T, Cp), (B * T, G * Cp), torch.bfloat16, B, T, g_a_col=0, g_out_col=Cp, **kw)
    xf = x.float()[..., :C].transpose(1, 2)
    for g in range(G):
        u = O.causal_conv1d(xf, cw[g].bfloat16().float(), cb[g], 2 ** g)
        if cond:
            gam = table[t_idx.long(), g, :C][:, :, None]
            bet = table[t_idx.long(), g, Cp:Cp + C][:, :, None]
            u = u * gam + bet
        want = (u.tanh() * u.sigmoid() + O.causal_conv1d(xf, rw[g].bfloat16().float(), rb[g])).transpose(1, 2)
        got = tc.view(B, T, G * Cp)[..., g * Cp:g * Cp + C]
        torch.testing.assert_close(got, want, rtol=2e-2, atol=3e-2)
        assert (tc.view(B, T, G * Cp)[..., g * Cp + C:(g + 1) * Cp] == 0).all()
    torch.testing.assert_close(tc, chk, rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("z", [16, 128])
def test_gemm_ddim_epilogue_matches_gemm_plus_ddim_step(z):
    """DN_EPI_DDIM (the sampler update inside the last GEMM's epilogue) against EPI_F32 + dn_ddim_step on the same inputs:
    the fp32 latent and both halves of the split-precision staging copy."""
    B, T, K, zp = 3, 333, 512, 128 if z > 64 else 64
    A = rnd(B * T, K, seed=95).half()
    W, b = rnd(z, K, seed=96, scale=0.05), rnd(z, seed=97)
    plan = packing.pack_linear(W.cpu(), b.cpu(), epi=_lib.EPI_F32, n_pad=z, fmt="f16").to(DEV)
    rows = torch.from_numpy(DDPMScheduler(200).ddim_rows()).to(DEV)
    t_idx = torch.tensor([57], dtype=torch.int32, device=DEV)
    x0 = rnd(B * T, z, seed=98)
    eh = torch.empty(B * T, z, device=DEV)
    plan.run(A, eh, B, T)
    x_ref, xb_ref = x0.clone(), torch.zeros(B * T, 2 * zp, dtype=torch.bfloat16, device=DEV)
    ops.ddim_step(x_ref, eh, rows, t_idx, 0, xb_ref, zp)
    for impl in (_lib.GEMM_SIMT_CHECK, _lib.GEMM_TCGEN05, _lib.GEMM_TCGEN05_2CTA):
        x, xb = x0.clone(), torch.zeros(B * T, 2 * zp, dtype=torch.bfloat16, device=DEV)
        plan.run(A, x, B, T, ddim=(rows, t_idx, xb, zp), impl=impl)
        torch.testing.assert_close(x, x_ref, rtol=1e-5, atol=1e-5)
        full = lambda p: p[:, :z].double() + p[:, zp:zp + z].double()
        assert (full(xb) - x.double()).abs().max() <= 2.0 ** -16 * x.abs().max()
        assert (xb[:, z:zp] == 0).all() and (xb[:, zp + z:] == 0).all()


def test_gemm_argmax_epilogue_matches_torch_argmax():
    """DN_EPI_ARGMAX + dn_argmax_combine (the unit head without materialised logits) == torch.argmax(logits) - 4 on the
    logits the same plan writes with EPI_F32, including ties (first index), NaN (greatest) and the excluded pad columns."""
    B, T, K, V = 2, 257, 768, 1004
    A = rnd(B * T, K, seed=100).bfloat16()
    W, b = rnd(V, K, seed=101, scale=0.05), rnd(V, seed=102)
    W[7] = W[3]; b[7] = b[3]                                  # exact tie between classes 3 and 7 in every row
    b[3] += 100.0; b[7] += 100.0                              # ... and they win: first index must be reported
    plan = packing.pack_linear(W.cpu(), b.cpu(), epi=_lib.EPI_F32, n_pad=1008).to(DEV)
    assert plan.n_tiles == 4
    A[5] = float("nan")                                       # a NaN row: every logit NaN -> index 0
    logits = torch.empty(B * T, 1008, device=DEV)
    plan.run(A, logits, B, T)
    want = torch.argmax(logits[:, :V], dim=-1) - 4
    assert (want[[0, 1, 2]] == 3 - 4).all() and want[5] == -4
    for impl in (_lib.GEMM_SIMT_CHECK, _lib.GEMM_TCGEN05, _lib.GEMM_TCGEN05_2CTA):
        parts = torch.full((B * T, 16), 7.0, device=DEV)
        plan.run(A, parts, B, T, argmax_classes=V, impl=impl)
        got = ops.argmax_combine(parts, 4)
        assert torch.equal(got, want), impl
    assert torch.equal(ops.argmax_units(logits, V, 4), want)


@pytest.mark.parametrize("B,T", [(5, 600), (7, 200), (3, 1000), (9, 136)])
def test_gemm_packed_rows_ragged_batches(B, T):
    """The config-4 shapes: a dilated causal conv (shifts up to 2 x 128 frames: taps must zero-fill at every utterance start,
    not read the previous utterance's tail), a WaveNet-style conditioned level and the split-precision form, over tiles that
    straddle utterances, against the SIMT checker and the per-utterance tiling."""
    Cin, N = 128, 272
    x = rnd(B, T, Cin, seed=110).bfloat16()
    W, b = rnd(N, Cin, 3, seed=111, scale=0.1), rnd(N, seed=112)
    for dil in (1, 16, 128):
        plan = packing.pack_conv3(W.cpu(), b.cpu(), dilation=dil).to(DEV)
        chk, tc = _gemm_case(plan, x.view(B * T, Cin), (B * T, N), torch.bfloat16, B, T)
        want = O.causal_conv1d(x.float().transpose(1, 2), W.bfloat16().float(), b, dil).transpose(1, 2).reshape(B * T, N)
        torch.testing.assert_close(tc, want, rtol=2e-2, atol=3e-2)
        torch.testing.assert_close(tc, chk, rtol=1e-2, atol=1e-2)
    # fp32 output with positions + lengths (per-row utterance parameters) and the in-place residual add (TMA reduce per chunk)
    lengths = torch.randint(1, T + 1, (B,), generator=torch.Generator().manual_seed(5)).to(torch.int32).to(DEV)
    pe = O.sinusoid_table(T + 1, 512).to(DEV).contiguous()
    Wl = rnd(512, Cin, seed=113, scale=0.1)
    plan = packing.pack_linear(Wl.cpu(), None, epi=_lib.EPI_F32).to(DEV)
    chk, tc = _gemm_case(plan, x.view(B * T, Cin), (B * T, 512), torch.float32, B, T, pe=pe, lengths=lengths)
    torch.testing.assert_close(tc, chk, rtol=1e-3, atol=1e-3)


def test_gemm_persistent_many_tiles():
    # > 148 tiles per launch: exercises the smem ring wrap-around and both TMEM accumulator stages
    B, T, K, N = 8, 1000, 512, 1536
    A = rnd(B * T, K, seed=90).bfloat16()
    W = rnd(N, K, seed=91, scale=0.05)
    plan = packing.pack_linear(W.cpu(), None).to(DEV)
    out = torch.empty(B * T, N, dtype=torch.bfloat16, device=DEV)
    plan.run(A, out, B, T)
    want = A.float() @ W.bfloat16().float().T
    torch.testing.assert_close(out.float(), want, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("M,K,bias,cond", [(128, 512, False, True), (1000, 512, False, True), (12000, 1365, True, True),
                                           (777, 1365, True, False), (40000, 512, True, True)])
def test_gemm_resid_norm_matches_unfused_pair(M, K, bias, cond):
    """dn_gemm_resid_norm (2-CTA cluster, DSMEM row-sum exchange) against EPI_RESID dn_gemm + dn_adarmsnorm: the
    residual stream bit-identical (same (acc + bias) + x order), the normalised bf16 operand within one bf16 ulp (the
    row sum of squares is formed in a different order).  Sizes cover one tile, ragged last tiles, > 74 clusters' worth
    of tiles (several tiles per cluster: ring wrap-around, both accumulators, the 4-deep partial-sum slots)."""
    W = rnd(512, K, seed=M + 1, scale=K ** -0.5)
    b = rnd(512, seed=M + 2) if bias else None
    plan = packing.pack_linear(W.cpu(), None if b is None else b.cpu(), epi=_lib.EPI_RESID, name="t").to(DEV)
    A = torch.zeros(M, plan.W.shape[1], dtype=torch.bfloat16, device=DEV)
    A[:, :K] = rnd(M, K, seed=M + 3).bfloat16()
    x0 = rnd(M, 512, seed=M + 4, scale=3.0)
    table = rnd(7, 1024, seed=M + 5)
    table[:, :512] += 1.0
    t_idx = torch.tensor([3], dtype=torch.int32, device=DEV)
    gp = rnd(512, seed=M + 6, scale=0.1) + 1
    x1, x2 = x0.clone(), x0.clone()
    h1 = torch.zeros(M, 512, dtype=torch.bfloat16, device=DEV)
    h2 = torch.zeros_like(h1)
    plan.run(A, x1, 1, M)
    if cond:
        ops.adarmsnorm(x1, h1, 1, M, None, table, 1024, t_idx, 0)
        ops.gemm_resid_norm(plan, A, x2, h2, None, table, 1024, t_idx)
    else:
        ops.adarmsnorm(x1, h1, 1, M, gp)
        ops.gemm_resid_norm(plan, A, x2, h2, gp)
    torch.cuda.synchronize()
    assert torch.equal(x1, x2)
    d = (h1.float() - h2.float()).abs()
    assert bool((d <= h1.float().abs() * 2 ** -7 + 1e-6).all()), f"max |dh| {d.max().item()}"
    # and the pair itself against fp32 torch (tolerance of a bf16-operand GEMM)
    want = x0 + A[:, :K].float() @ W.bfloat16().float().T + (0 if b is None else b)
    torch.testing.assert_close(x2, want, rtol=2e-2, atol=2e-2)


def test_kmeans_quantizer_matches_sklearn_and_oracle():
    """dn_split_bf16x3 + dn_gemm (K = 3 x 768) + dn_argmax_units vs scikit-learn's labels (golden) and the float64 oracle.
    Bar: identical labels wherever the float64 top-2 squared-distance gap exceeds 1e-3 (near-ties below that are within
    fp32 rounding of scikit-learn's own float32 arithmetic); >= 99.9 % overall."""
    import os
    from diffnorm_b200.kmeans import KMeansQuantizer
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmeans_predict.npz"))
    centers, feats = O.kmeans_case(int(g["seed"]), int(g["K"]), int(g["D"]), int(g["N"]))
    q = KMeansQuantizer(centers)
    got = q.predict(torch.from_numpy(feats).cuda()).cpu().numpy()
    want = g["labels"].astype(np.int64)
    c, x = centers.astype(np.float64), feats.astype(np.float64)
    d = (c * c).sum(1)[None, :] - 2.0 * x @ c.T
    part = np.partition(d, 1, axis=1)
    confident = (part[:, 1] - part[:, 0]) > 1e-3
    print(f"[parity] kmeans: agreement {np.mean(got == want):.5f} overall, {np.mean(got[confident] == want[confident]):.5f} on "
          f"{confident.mean():.4f} confident frames")
    assert (got[confident] == want[confident]).all() and np.mean(got == want) >= 0.999
    parts = q.predict_many([feats[:700], feats[700:701], feats[701:]], max_rows=1024)
    assert np.array_equal(np.concatenate(parts), got)
