"""Operand-format kernels (B200 only): fp16 weights, fp16 outputs and the split-precision (hi | lo bf16 pair) path that the
VAE encoder / decoder run on, each against a float64 statement of the same op and against the SIMT checker kernel.

Stated tolerances: a split-precision contraction carries ~2^-17 per operand, so results must sit within 1e-4 of float64
relative to the output scale (a bf16 contraction sits at ~3e-3); tensor-core and checker kernels must agree to fp32
accumulation-order noise."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from diffnorm_b200 import _lib, ops, packing  # noqa: E402
from oracle import diffnorm_oracle as O  # noqa: E402

DEV = "cuda"
bf16, f16, f32 = torch.bfloat16, torch.float16, torch.float32


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def split(x, width=None):
    """fp32 [M, C] -> [M, 2 * width] split pair through dn_cast_split."""
    M, C = x.shape
    w = width or C
    out = torch.full((M, 2 * w), 7.0, dtype=bf16, device=DEV)
    ops.cast_split(x.contiguous(), out, w)
    return out


def join(pair):
    w = pair.shape[-1] // 2
    return pair[..., :w].double() + pair[..., w:].double()


def rel(got, want):
    return float((got.double() - want.double()).abs().max() / want.double().abs().max())


def run_all(plan, A, out_shape, dt, B, T, **kw):
    outs = []
    for impl in (_lib.GEMM_SIMT_CHECK, _lib.GEMM_TCGEN05, _lib.GEMM_TCGEN05_2CTA):
        torch.manual_seed(0)
        out = torch.randn(out_shape, device=DEV).to(dt).contiguous()
        plan.run(A, out, B, T, impl=impl, **kw)
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[1], outs[2]), "cta_group::2 kernel differs from the single-CTA kernel"
    return outs[0], outs[1]


def test_cast_split_layout_and_precision():
    x = rnd(70, 20, seed=1, scale=3.0)
    p = split(x, 64)
    assert p.shape == (70, 128)
    assert torch.equal(p[:, :20].float(), x.bfloat16().float())
    assert (p[:, 20:64] == 0).all() and (p[:, 84:] == 0).all()
    assert rel(join(p)[:, :20], x) < 2 ** -16
    q = torch.full((70, 64), 7.0, dtype=bf16, device=DEV)
    ops.cast_split(x, q, 0)   # lo_col 0 = plain cast + zero pad
    assert torch.equal(q[:, :20].float(), x.bfloat16().float()) and (q[:, 20:] == 0).all()


def test_gemm_split_linear_is_fp32_grade():
    B, T, K, N = 2, 200, 192, 272
    A, W, b = rnd(B * T, K, seed=2), rnd(N, K, seed=3, scale=0.1), rnd(N, seed=4)
    want = A.double() @ W.double().T + b.double()
    plan = packing.pack_linear(W.cpu(), b.cpu(), epi=_lib.EPI_F32, fmt="split").to(DEV)
    chk, tc = run_all(plan, split(A), (B * T, N), f32, B, T)
    assert rel(tc, want) < 1e-4, rel(tc, want)
    assert rel(tc, chk) < 5e-5
    # the same product with bf16 operands is ~50x worse: the test would not pass by accident
    plain = packing.pack_linear(W.cpu(), b.cpu(), epi=_lib.EPI_F32).to(DEV)
    out = torch.empty(B * T, N, device=DEV)
    plain.run(A.bfloat16(), out, B, T)
    assert rel(out, want) > 1e-3
    # split output: hi + lo reproduces the fp32 result, RESID accumulates it
    plan_b = packing.pack_linear(W[:256].cpu(), b[:256].cpu(), epi=_lib.EPI_BF16, fmt="split").to(DEV)
    chk, tc = run_all(plan_b, split(A), (B * T, 512), bf16, B, T, out_split=True)
    assert rel(join(tc), want[:, :256]) < 1e-4
    assert torch.equal(tc[:, :256], chk[:, :256]) or rel(join(tc), join(chk)) < 5e-5


@pytest.mark.parametrize("dil", [1, 4])
def test_gemm_split_causal_conv(dil):
    B, T, Cin, N = 2, 300, 128, 192
    x = rnd(B, T, Cin, seed=5)
    W, b = rnd(N, Cin, 3, seed=6, scale=0.1), rnd(N, seed=7)
    plan = packing.pack_conv3(W.cpu(), b.cpu(), dilation=dil, fmt="split").to(DEV)
    assert len(plan.segs) == 3
    chk, tc = run_all(plan, split(x.view(B * T, Cin)), (B * T, 2 * N), bf16, B, T, out_split=True)
    want = O.causal_conv1d(x.double().transpose(1, 2), W.double(), b.double(), dil).transpose(1, 2).reshape(B * T, N)
    assert rel(join(tc), want) < 1e-4, rel(join(tc), want)
    assert rel(join(tc), join(chk)) < 5e-5


def test_gemm_split_geglu_precise():
    B, T, K, inner = 2, 130, 128, 256
    A = rnd(B * T, K, seed=8)
    W, b = rnd(2 * inner, K, seed=9, scale=0.1), rnd(2 * inner, seed=10)
    plan = packing.pack_geglu(W.cpu(), b.cpu(), fmt="split").to(DEV)
    chk, tc = run_all(plan, split(A), (B * T, 2 * inner), bf16, B, T, out_split=True)
    h = A.double() @ W.double().T + b.double()
    want = torch.nn.functional.gelu(h[:, inner:]) * h[:, :inner]
    assert rel(join(tc), want) < 1e-4, rel(join(tc), want)
    assert rel(join(tc), join(chk)) < 5e-5


def test_gemm_split_wavenet_level_precise():
    B, T, C, G = 2, 260, 128, 3
    x = rnd(B, T, C, seed=11)
    cw = [rnd(C, C, 3, seed=12 + g, scale=0.08) for g in range(G)]
    cb = [rnd(C, seed=20 + g) for g in range(G)]
    rw = [rnd(C, C, 1, seed=30 + g, scale=0.08) for g in range(G)]
    rb = [rnd(C, seed=40 + g) for g in range(G)]
    plan = packing.pack_wavenet_level([w.cpu() for w in cw], [w.cpu() for w in cb], [w.cpu() for w in rw],
                                      [w.cpu() for w in rb], C, fmt="split").to(DEV)
    chk, tc = run_all(plan, split(x.view(B * T, C)), (B * T, 2 * G * C), bf16, B, T, g_a_col=0, g_out_col=C, out_split=True)
    got = join(tc).view(B, T, G * C)
    xd = x.double().transpose(1, 2)
    for g in range(G):
        u = O.causal_conv1d(xd, cw[g].double(), cb[g].double(), 2 ** g)
        want = (u.tanh() * u.sigmoid() + O.causal_conv1d(xd, rw[g].double(), rb[g].double())).transpose(1, 2)
        assert rel(got[..., g * C:(g + 1) * C], want) < 1e-4
    assert rel(join(tc), join(chk)) < 5e-5
    # second level of a stack: the A operand is the grouped split buffer (group g at columns g*C, lo half at G*C + g*C)
    chk2, tc2 = run_all(plan, tc, (B * T, 2 * G * C), bf16, B, T, g_a_col=C, g_out_col=C, out_split=True)
    y = got
    for g in range(G):
        yg = y[..., g * C:(g + 1) * C].transpose(1, 2)
        u = O.causal_conv1d(yg, cw[g].double(), cb[g].double(), 2 ** g)
        want = (u.tanh() * u.sigmoid() + O.causal_conv1d(yg, rw[g].double(), rb[g].double())).transpose(1, 2)
        assert rel(join(tc2).view(B, T, G * C)[..., g * C:(g + 1) * C], want) < 1e-4
    assert rel(join(tc2), join(chk2)) < 5e-5


def test_gemm_fp16_operands_and_fp16_output():
    """The sampler loop's format: fp16 activations x fp16 weights (one format per MMA: a bf16 x fp16 mix is rejected)."""
    B, T, K, N = 2, 200, 192, 256
    A32 = rnd(B * T, K, seed=50)
    A = A32.half()
    W, b = rnd(N, K, seed=51, scale=0.1), rnd(N, seed=52)
    plan = packing.pack_linear(W.cpu(), b.cpu(), epi=_lib.EPI_F32, fmt="f16").to(DEV)
    assert plan.W.dtype == f16
    chk, tc = run_all(plan, A, (B * T, N), f32, B, T)
    want = A.double() @ W.half().double().T + b.double()     # exact product of the rounded operands
    assert rel(tc, want) < 1e-5, rel(tc, want)
    assert rel(tc, chk) < 5e-5
    with pytest.raises(ValueError):
        plan.run(A32.bfloat16(), torch.empty(B * T, N, device=DEV), B, T)
    # and it is ~8x closer to the un-rounded product than the bf16 plan is (what the sampler loop gains)
    full = A32.double() @ W.double().T + b.double()
    plain = packing.pack_linear(W.cpu(), b.cpu(), epi=_lib.EPI_F32).to(DEV)
    out = torch.empty(B * T, N, device=DEV)
    plain.run(A32.bfloat16(), out, B, T)
    assert rel(tc, full) < 0.25 * rel(out, full)
    # fp16 outputs of the three 16-bit epilogues
    pb = packing.pack_linear(W.cpu(), b.cpu(), epi=_lib.EPI_BF16, fmt="f16").to(DEV)
    chk, tc = run_all(pb, A, (B * T, N), f16, B, T)
    assert rel(tc, want) < 1e-3 and (tc.float() - chk.float()).abs().max() <= 2e-3 * want.abs().max()
    big = packing.pack_linear((W * 1e4).cpu(), None, epi=_lib.EPI_BF16, fmt="f16").to(DEV)
    o = torch.empty(B * T, N, dtype=f16, device=DEV)
    big.run((A32 * 100).half(), o, B, T)
    assert torch.isfinite(o.float()).all() and o.float().abs().max() == 65504.0     # saturates, never inf
    Wg, bg = rnd(512, K, seed=54, scale=0.1), rnd(512, seed=55)
    pg = packing.pack_geglu(Wg.cpu(), bg.cpu(), fmt="f16").to(DEV)
    chk, tc = run_all(pg, A, (B * T, 256), f16, B, T)
    h = A.double() @ Wg.half().double().T + bg.double()
    wantg = torch.nn.functional.gelu(h[:, 256:]) * h[:, :256]
    assert rel(tc, wantg) < 2e-3 and rel(tc, chk) < 2e-3
    # fp16 output of a split-precision plan (the precise decoder's q, k, v)
    plan2 = packing.pack_linear(W.cpu(), b.cpu(), epi=_lib.EPI_BF16, fmt="split").to(DEV)
    chk, tc = run_all(plan2, split(A32), (B * T, N), f16, B, T, out_f16=True)
    want = A32.double() @ W.double().T + b.double()
    assert rel(tc, want) < 1e-3      # one fp16 rounding of the output
    assert (tc.float() - chk.float()).abs().max() <= 2e-3 * want.abs().max()


def test_attention_dh64_fp16():
    B, T, H, dh = 2, 300, 8, 64
    qkv = rnd(B, T, 3 * H * dh, seed=56).half()
    lengths = torch.tensor([300, 131], dtype=torch.int32, device=DEV)
    out = torch.empty(B, T, H * dh, dtype=f16, device=DEV)
    ops.attention(qkv, out, lengths, B, T, H, dh)
    q, k, v = (t.double().view(B, T, H, dh).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * dh ** -0.5
    mask = torch.arange(T, device=DEV)[None] < lengths[:, None]
    sim = sim.masked_fill(~mask[:, None, None, :], -torch.finfo(torch.float64).max)
    want = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), v).transpose(1, 2).reshape(B, T, H * dh)
    assert rel(out, want) < 3e-3, rel(out, want)     # fp16 P and output; the bf16 kernel sits at ~1e-2


def test_stochastic_weight_rerounding_is_unbiased_and_step_dependent():
    """dn_sround_bf16: every output is one of the two bf16 neighbours of the fp32 master, a step's rounding is a
    deterministic function of (seed, step), differs between steps, and averages to the master (E[dst] = src)."""
    n = 1 << 16
    w = rnd(n, seed=90, scale=0.05)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    dst = torch.empty(n, dtype=bf16, device=DEV)
    lo = w.view(torch.int32).bitwise_and(-65536).view(torch.float32)           # truncated toward zero
    hi = (w.view(torch.int32).bitwise_and(-65536) + 65536).view(torch.float32)    # next bf16 away from zero
    acc = torch.zeros(n, dtype=torch.float64, device=DEV)
    outs = []
    for t in range(64):
        step.fill_(t)
        ops.sround_bf16(w, dst, 1234, step)
        v = dst.float()
        assert bool(((v == lo) | (v == hi)).all())
        acc += v.double()
        outs.append(v.clone())
    assert not torch.equal(outs[0], outs[1])
    step.fill_(0)
    ops.sround_bf16(w, dst, 1234, step)
    assert torch.equal(dst.float(), outs[0])                                  # deterministic given (seed, step)
    ops.sround_bf16(w, dst, 99, step)
    assert not torch.equal(dst.float(), outs[0])
    mean_err = (acc / 64 - w.double()).abs().mean() / w.abs().mean()
    rn_err = (w.bfloat16().double() - w.double()).abs().mean() / w.abs().mean()
    assert mean_err < 0.3 * rn_err, (float(mean_err), float(rn_err))          # 64 independent draws: ~1/8 of one rounding
    up = (torch.stack(outs) == hi).double().mean(0)                           # P(round up) = fractional position
    frac = ((w.double() - lo.double()) / (hi.double() - lo.double()))
    assert float((up - frac).abs().mean()) < 0.06


@pytest.mark.parametrize("C", [512, 768])
def test_adarmsnorm_split(C):
    B, T = 3, 41
    x = rnd(B * T, C, seed=60, scale=2.0)
    gp = rnd(C, seed=61) + 1
    out = torch.empty(B * T, 2 * C, dtype=bf16, device=DEV)
    ops.adarmsnorm(x, out, B, T, gp, split=True)
    want = O.rmsnorm(x.double().view(B, T, C), gp.double()).view(B * T, C)
    assert rel(join(out), want) < 1e-4
    ref = torch.empty(B * T, C, dtype=bf16, device=DEV)
    ops.adarmsnorm(x, ref, B, T, gp)
    assert torch.equal(out[:, :C], ref)          # hi half = the plain kernel's output


def test_latent_updates_write_split_staging():
    rows, z, zp = 333, 16, 64
    zl, eps = rnd(rows, z, seed=70), rnd(rows, z, seed=71)
    x = torch.empty(rows, z, device=DEV)
    xb = torch.full((rows, 2 * zp), 5.0, dtype=bf16, device=DEV)
    ops.q_sample(zl, eps, 0.6, 0.8, x, xb, zp)
    assert rel(join(xb)[:, :z], x) < 2 ** -16
    assert (xb[:, z:zp] == 0).all() and (xb[:, zp + z:] == 0).all()
    table = torch.tensor([[0.7, 0.71, 0.8, 0.6, 1.4, 1.0, 0, 0]], device=DEV)
    t_idx = torch.zeros(1, dtype=torch.int32, device=DEV)
    eh = rnd(rows, z, seed=72)
    x_ref = x.clone()
    ops.ddim_step(x_ref, eh, table, t_idx, 0)
    xb.fill_(5.0)
    ops.ddim_step(x, eh, table, t_idx, 0, xb, zp)
    assert torch.equal(x, x_ref)
    assert rel(join(xb)[:, :z], x) < 2 ** -16
    assert (xb[:, z:zp] == 0).all() and (xb[:, zp + z:] == 0).all()
    noise = rnd(rows, z, seed=73)
    xb.fill_(5.0)
    ops.ddpm_step(x, eh, noise, table, t_idx, xb, zp)
    assert rel(join(xb)[:, :z], x) < 2 ** -16


@pytest.mark.parametrize("T,lens", [(77, [77, 1]), (300, [300, 171])])
def test_attention_fp16_split_output(T, lens):
    B, H, dh = 2, 8, 96
    qkv32 = rnd(B, T, 3 * H * dh, seed=80)
    qkv = qkv32.half()
    lengths = torch.tensor(lens, dtype=torch.int32, device=DEV)
    out = torch.empty(B, T, 2 * H * dh, dtype=bf16, device=DEV)
    ops.attention(qkv, out, lengths, B, T, H, dh, out_split=True)
    q, k, v = (t.double().view(B, T, H, dh).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * dh ** -0.5
    mask = torch.arange(T, device=DEV)[None] < lengths[:, None]
    sim = sim.masked_fill(~mask[:, None, None, :], -torch.finfo(torch.float64).max)
    want = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), v).transpose(1, 2).reshape(B, T, H * dh)
    got = join(out)
    # fp16 probabilities (2^-11) against exact q, k, v: well inside 2e-3 of the output scale; bf16 would be ~1e-2
    assert rel(got, want) < 2e-3, rel(got, want)
