// dn_gemm: bf16 x bf16 -> fp32 GEMM / implicit causal dilated convolution on tcgen05 tensor cores.
//
// Replaces every nn.Linear / CausalConv1d on the reference path (LM:476-488, :509-511, :887-903, :930-932;
// SURVEY §2.3 G1-G8).  One persistent, warp-specialised kernel:
//   warp 0      TMA producer   A tile  = box {64 ch, 128 frames, 1 utt} of a 3-D map [C, T, B]; a conv tap is the
//                              same box shifted by -shift frames, TMA zero-fills t < 0 (= causal left padding)
//                              W tile  = box {64, 256 | 128} of the packed K-major weight matrix
//   warp 1      MMA issuer     tcgen05.mma cta_group::1 kind::f16, M = 128, N <= 256, K = 16 per instruction,
//                              fp32 accumulators in TMEM (2 x 256 columns, double buffered against the epilogue)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue       tcgen05.ld 32x32b (one accumulator row per thread) -> fused epilogue -> global
// Epilogues: +bias -> bf16 | fp32 (+ sinusoidal positions) | in-place fp32 residual add | GEGLU | WaveNet
// FiLM + tanh*sigmoid gate + residual branch (second accumulator half).
#include "common.cuh"

namespace dn {

unsigned long long g_launch_count = 0;

constexpr int BM = 128;         // frames per tile (UMMA M)
constexpr int BK = 64;          // bf16 per K block = one 128-byte swizzle atom
constexpr int WT = 256;         // packed weight rows per N tile (max UMMA N)
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = WT * BK * 2;
// smem ring: single CTA = 4 stages of (A 16 KB + W 32 KB); CTA pair = 6 stages of (A 16 KB + half of W 16 KB)
template <int CTAS> struct Ring {
    static constexpr int STAGES = CTAS == 2 ? 6 : 4;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES / CTAS;
};
constexpr int ACC_COLS = 256;   // TMEM columns per accumulator stage
constexpr int GEMM_THREADS = 384;
constexpr int EPI_WARPS = 8;
constexpr int UNIT_BYTES = 128 * 128;  // epilogue staging unit: 128 rows x 128 B (one TMA store box)
constexpr int BAR_BYTES = 2048;      // mbarriers (first 512 B) + per-tile epilogue parameter staging (GEGLU bias, 1 KB)
constexpr int GEMM_SMEM = 4 * (A_BYTES + B_BYTES) + BAR_BYTES + 2 * UNIT_BYTES + 1024 /*align slack*/;  // both ring forms = 192 KB

struct TileCoord {
    int g, b, t0, n;
    int chunk0;    // packed rows: first row chunk of this CTA's 128 rows
    bool contig;   // the 128 rows lie inside one utterance: one box at (b, t0) moves them
};
// M tiling of a launch.  Plain: per_t tiles per utterance, a tile never leaves its utterance.  Packed rows (row_chunk = ch):
// per_t chunks per utterance, tiles take bm / ch consecutive chunks of the batch's chunk sequence.
struct Tiling {
    int m_tiles, per_t, ch;
};
__host__ __device__ __forceinline__ Tiling make_tiling(const dn_gemm_desc& p, int bm) {
    Tiling g;
    g.ch = p.row_chunk;
    if (g.ch) {
        g.per_t = (p.T + g.ch - 1) / g.ch;
        g.m_tiles = (int)(((long long)p.B * g.per_t * g.ch + bm - 1) / bm);
    } else {
        g.per_t = (p.T + bm - 1) / bm;
        g.m_tiles = p.B * g.per_t;
    }
    return g;
}
// bm = rows per tile (128, or 256 for a CTA pair); row_off = this CTA's offset inside the tile
__host__ __device__ __forceinline__ TileCoord decode_tile(const dn_gemm_desc& p, int tile, const Tiling& tg, int bm = BM, int row_off = 0) {
    TileCoord c;
    c.n = tile % p.n_tiles;
    int r = tile / p.n_tiles;
    int m = r % tg.m_tiles;
    c.g = r / tg.m_tiles;
    if (tg.ch) {
        c.chunk0 = (m * bm + row_off) / tg.ch;
        c.b = c.chunk0 / tg.per_t;           // may be >= B for the tail of the last tile: loads zero-fill, stores are clipped
        const int r0 = c.chunk0 - c.b * tg.per_t;
        c.t0 = r0 * tg.ch;
        c.contig = r0 + BM / tg.ch <= tg.per_t;
    } else {
        c.b = m / tg.per_t;
        c.t0 = (m % tg.per_t) * bm + row_off;
        c.chunk0 = 0;
        c.contig = true;
    }
    return c;
}

__device__ __forceinline__ const float* gb_row(const dn_gemm_desc& p, int b, int g) {
    if (!p.gb) return nullptr;
    int t = p.t_idx ? p.t_idx[(long long)b * p.t_idx_stride] : 0;
    return p.gb + (long long)t * p.gb_t_stride + (long long)g * p.g_gb;
}

// CTAS = 2: the kernel runs as a cluster of two CTAs (one TPC) on M = 256 tiles with tcgen05.mma.cta_group::2: each CTA
// loads its own 128 A rows and HALF of the W tile's rows, the leader CTA issues the MMAs for both, completion is
// multicast to both CTAs' barriers, and each CTA runs the unchanged epilogue on its own 128 accumulator rows.  Per flop
// the pair reads a third less shared memory and fetches each W tile once instead of twice.
// MODE 0: 16-bit outputs are bf16; MODE 1: split-precision output (hi | lo bf16 pairs, full-precision erf / tanh / exp in
// the GEGLU / gate epilogues); MODE 2: fp16 output (one saturating F2FP per pair, same cost as the bf16 pack).
template <int EPI, int CTAS, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmW128, const __grid_constant__ CUtensorMap tmW64,
               const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAC,
               const __grid_constant__ CUtensorMap tmOutC, const dn_gemm_desc p) {
    constexpr int STAGES = Ring<CTAS>::STAGES;
    constexpr int STAGE_BYTES = Ring<CTAS>::STAGE_BYTES;
    static_assert(STAGES * STAGE_BYTES == 4 * (A_BYTES + B_BYTES), "both ring forms use the same 192 KB");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space: LDS / STS
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    pdl_trigger();    // the next kernel may be scheduled onto SMs as this grid's CTAs retire
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    constexpr int TBM = BM * CTAS;                 // rows per tile
    const int row_off = (int)rank * BM;
    const Tiling tg = make_tiling(p, TBM);
    const int total = p.groups * tg.m_tiles * p.n_tiles;
    const int nbox = tg.ch ? BM / tg.ch : 1;       // boxes per A tile / output unit of a tile that straddles utterances
    const int rows_pg = p.groups > 1 ? p.g_w_row : p.w_rows;
    const int tile0 = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmW128);
        tma_prefetch_desc(&tmW64);
        tma_prefetch_desc(&tmOut);
        if (tg.ch) {
            tma_prefetch_desc(&tmAC);
            tma_prefetch_desc(&tmOutC);
        }
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], EPI_WARPS * CTAS);   // pair: the leader's MMA waits for both CTAs' epilogues
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CTAS == 2) {
            tmem_alloc2(tmem_slot, 512);
            tmem_relinquish2();
        } else {
            tmem_alloc(tmem_slot, 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();       // everything above touched only this CTA's smem / TMEM; inputs are read from here on

    if (warp == 0) {
        // lane 0 produces; with packed rows the other lanes stay in the loop to issue one chunk box each for the tiles
        // that straddle an utterance boundary
        if (lane == 0 || tg.ch) {
            // ---------------------------------------------------------------- TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < total; tile += tile_step) {
                TileCoord c = decode_tile(p, tile, tg, TBM, row_off);
                const int d = p.dilation << (p.dilation_shl_group ? c.g : 0);
                int n_full = rows_pg - c.n * WT;
                n_full = n_full > WT ? WT : n_full;
                const bool boxes = !c.contig;          // warp-uniform
                int lb = 0, lt0 = 0;                   // this lane's chunk of a straddling tile
                if (boxes && lane < nbox) {
                    const int chn = c.chunk0 + lane;
                    lb = chn / tg.per_t;
                    lt0 = (chn - lb * tg.per_t) * tg.ch;
                }
                for (int s = 0; s < p.num_segs; ++s) {
                    const dn_gemm_seg sg = p.seg[s];
                    const bool half_w = sg.n_mma == 128;
                    const int nrows = sg.n_mma ? sg.n_mma : n_full;   // W rows this segment multiplies
                    for (int kb = 0; kb < sg.k_blocks; ++kb) {
                        uint8_t* sa = smem + stage * STAGE_BYTES;
                        uint8_t* sb = sa + A_BYTES;
                        const int acol = sg.a_col0 + c.g * p.g_a_col + kb * BK;
                        if (lane == 0) {
                            mbar_wait(&empty[stage], phase ^ 1);
                            if (CTAS == 1) {
                                mbar_expect_tx(&full[stage], A_BYTES + (half_w ? B_BYTES / 2 : B_BYTES));
                                if (!boxes) tma_load_3d(&tmA, &full[stage], sa, acol, c.t0 - sg.shift_mul * d, c.b);
                                tma_load_2d(half_w ? &tmW128 : &tmW, &full[stage], sb, sg.w_k0 + kb * BK,
                                            c.g * p.g_w_row + c.n * WT);
                            } else {
                                // each CTA brings nrows / 2 W rows (box of 128 or 64 rows; rows past the half are not read)
                                const int half = nrows >> 1;
                                const bool box128 = half > 64;
                                const uint32_t bytes = A_BYTES + (box128 ? 128 : 64) * BK * 2;
                                if (leader) mbar_expect_tx(&full[stage], 2 * bytes);
                                if (!boxes) tma2_load_3d(&tmA, &full[stage], sa, acol, c.t0 - sg.shift_mul * d, c.b);
                                tma2_load_2d(box128 ? &tmW128 : &tmW64, &full[stage], sb, sg.w_k0 + kb * BK,
                                             c.g * p.g_w_row + c.n * WT + (int)rank * half);
                            }
                        }
                        if (boxes) {
                            __syncwarp();      // lane 0 has seen the slot empty
                            if (lane < nbox) {
                                uint8_t* dst = sa + lane * tg.ch * (BK * 2);
                                if (CTAS == 1) tma_load_3d(&tmAC, &full[stage], dst, acol, lt0 - sg.shift_mul * d, lb);
                                else tma2_load_3d(&tmAC, &full[stage], dst, acol, lt0 - sg.shift_mul * d, lb);
                            }
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // ---------------------------------------------------------------- MMA issuer (one thread; pair: leader CTA)
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = tile0; tile < total; tile += tile_step) {
                TileCoord c = decode_tile(p, tile, tg, TBM, row_off);
                int n_full = rows_pg - c.n * WT;
                n_full = n_full > WT ? WT : n_full;
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * ACC_COLS;
                uint32_t acc = 0;
                for (int s = 0; s < p.num_segs; ++s) {
                    const dn_gemm_seg sg = p.seg[s];
                    const uint32_t nrows = sg.n_mma ? sg.n_mma : n_full;
                    const uint32_t idesc = umma_idesc_16(CTAS == 2 ? 256u : 128u, nrows, p.a_fmt, p.w_fmt);
                    for (int kb = 0; kb < sg.k_blocks; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t da = umma_desc_sw128(sa);
                        const uint64_t db = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // advance 16 bf16 = 32 B inside the 128-B swizzle atom: +2 in the (addr >> 4) field
                            if (CTAS == 2) umma2_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, acc);
                            else umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, acc);
                            acc = 1;
                        }
                        // smem slot reusable (in both CTAs of a pair) once these MMAs have read it
                        if (CTAS == 2) umma2_commit_mc(&empty[stage]); else umma_commit(&empty[stage]);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
                if (CTAS == 2) umma2_commit_mc(&tfull[as]); else umma_commit(&tfull[as]);  // accumulator complete -> epilogue(s)
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // -------------------------------------------------------------------- epilogue (2 warpgroups x 4 warps)
        // Each thread owns one accumulator row (tcgen05.ld 32x32b).  Results are staged in a 16 KB smem "unit"
        // (128 rows x 128 B, 128B-swizzled = the TMA box layout) per warpgroup and written with ONE TMA store
        // (EPI_RESID: TMA reduce-add into the fp32 residual stream), so global writes are fully coalesced,
        // asynchronous, and clipped at the tensor edge (t >= T, col >= n_out) by the hardware.
        const int q = warp & 3;            // TMEM lane quadrant this warp may access
        const int half = (warp - 4) >> 2;  // warpgroup: which units of the tile it handles
        const int row = q * 32 + lane;
        const bool issuer = (warp == 4 + 4 * half) && lane == 0;
        uint8_t* stage_buf = smem + STAGES * STAGE_BYTES + BAR_BYTES + half * UNIT_BYTES;
        uint8_t* srow = stage_buf + row * 128;
        const int sw = row & 7;
        const int bar_id = 1 + half;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = tile0; tile < total; tile += tile_step) {
            TileCoord c = decode_tile(p, tile, tg, TBM, row_off);
            // this thread's row: (utterance wb, frame t); a straddling tile maps every chunk on its own
            int wb = c.b, t = c.t0 + row;
            if (!c.contig) {
                const int chn = c.chunk0 + row / tg.ch;
                wb = chn / tg.per_t;
                t = (chn - wb * tg.per_t) * tg.ch + row % tg.ch;
            }
            const bool row_ok = t < p.T && wb < p.B;
            const int pb_ = wb < p.B ? wb : p.B - 1;     // utterance whose per-utterance parameters this row reads
            // one TMA box per unit, or one per chunk when the tile straddles utterances (issuer thread only)
            auto store_unit = [&](const uint8_t* src, int ocol, bool reduce) {
                if (c.contig) {
                    if (reduce) tma_reduce_add_3d(&tmOut, src, ocol, c.t0, c.b);
                    else tma_store_3d(&tmOut, src, ocol, c.t0, c.b);
                } else {
                    for (int i = 0; i < nbox; ++i) {
                        const int chn = c.chunk0 + i;
                        const int bb = chn / tg.per_t, tt = (chn - bb * tg.per_t) * tg.ch;
                        if (reduce) tma_reduce_add_3d(&tmOutC, src + i * tg.ch * 128, ocol, tt, bb);
                        else tma_store_3d(&tmOutC, src + i * tg.ch * 128, ocol, tt, bb);
                    }
                }
            };
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + as * ACC_COLS + ((uint32_t)(q * 32) << 16);
            const int ocol0 = c.g * p.g_out_col;

            if constexpr (EPI == DN_EPI_DDIM) {
                // eps_hat stays in registers: x <- sqrt(ab_prev) x0 + sqrt(1 - ab_prev) eps~ with x0, eps~ re-derived from x and
                // eps_hat exactly as the reference does (safe_div clamps 1e-10), then the fp32 state and its split-precision
                // staging copy are written straight from the thread that owns the row (rows are contiguous: coalesced)
                if (half == 0) {   // warp-uniform: tcgen05.ld is warp-collective, rows past T only skip the global accesses
                    const float* cf = p.coef + (long long)(p.t_idx ? p.t_idx[0] : 0) * 8;
                    const float c0 = cf[0], c1 = cf[1], c2 = cf[2], c3 = cf[3];
                    const float d0 = fmaxf(c0, 1e-10f), d1 = fmaxf(c1, 1e-10f);
                    const long long r = row_ok ? (long long)wb * p.T + t : 0;
                    float* xr = reinterpret_cast<float*>(p.out) + r * p.ldo;
                    __nv_bfloat16* sr = reinterpret_cast<__nv_bfloat16*>(p.aux) + r * p.aux_ld;
                    for (int cc = 0; cc < p.n_out; cc += 16) {
                        float e[16];
                        tmem_ld16(taddr + cc, e);
                        tmem_ld_wait();
                        if (!row_ok) continue;
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 xv = *reinterpret_cast<const float4*>(xr + cc + i);
                            const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + cc + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, bs[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                            for (int q2 = 0; q2 < 4; ++q2) {
                                const float eh = e[i + q2] + bs[q2];
                                const float x0 = (xs[q2] - c1 * eh) / d0;
                                const float pn = (xs[q2] - c0 * x0) / d1;
                                v[i + q2] = x0 * c2 + c3 * pn;
                            }
                            *reinterpret_cast<float4*>(xr + cc + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        }
#pragma unroll
                        for (int i = 0; i < 16; i += 8) {
                            uint32_t hh[4], ll[4];
#pragma unroll
                            for (int q2 = 0; q2 < 4; ++q2) split_bf16(v[i + 2 * q2], v[i + 2 * q2 + 1], hh[q2], ll[q2]);
                            *reinterpret_cast<uint4*>(sr + cc + i) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                            if (p.aux_lo_col) *reinterpret_cast<uint4*>(sr + p.aux_lo_col + cc + i) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                        }
                    }
                }
            } else if constexpr (EPI == DN_EPI_ARGMAX) {
                // this warpgroup's 128 accumulator columns of the row -> one (max, first index) partial; NaN is greatest and
                // the first occurrence wins (torch.argmax); columns are scanned in ascending order so a strict > keeps it
                const int c0 = c.n * WT + half * 128;
                float bv = -INFINITY;
                int bi = 0x7fffffff;
                for (int sub = 0; sub < 4; ++sub) {
                    const int cs = c0 + sub * 32;
                    if (cs >= p.n_classes) break;
                    float v[32];
                    tmem_ld32(taddr + half * 128 + sub * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int cc = cs + i;
                        if (cc < p.n_classes) {
                            const float x = v[i] + (p.bias ? __ldg(p.bias + cc) : 0.f);
                            const bool xn = x != x, bn = bv != bv;
                            if (bi == 0x7fffffff || (xn && !bn) || (!bn && x > bv)) { bv = x; bi = cc; }
                        }
                    }
                }
                if (row_ok) {
                    float* o = reinterpret_cast<float*>(p.out) + ((long long)wb * p.T + t) * p.ldo + (c.n * 2 + half) * 2;
                    *reinterpret_cast<float2*>(o) = make_float2(bv, __int_as_float(bi));
                }
            } else if constexpr (EPI == DN_EPI_BF16 || EPI == DN_EPI_F32 || EPI == DN_EPI_RESID) {
                constexpr int UCOLS = (EPI == DN_EPI_BF16) ? 64 : 32;   // columns per 16 KB unit
                const float* bias = p.bias ? p.bias + c.g * p.g_bias : nullptr;
                long long pe_row = -1;
                if (EPI == DN_EPI_F32 && p.pe && row_ok) {
                    int pos = t + 1;
                    if (p.lengths && t >= p.lengths[wb]) pos = 0;
                    pe_row = (long long)pos * p.n_out;
                }
                constexpr int npass = (EPI == DN_EPI_BF16 && MODE == 1) ? 2 : 1;   // split output: hi pass, then lo = v - hi
                for (int u = half; u < WT / UCOLS; u += 2) {
                    const int col = c.n * WT + u * UCOLS;
                    if (col >= p.n_out) break;
                    for (int pass = 0; pass < npass; ++pass) {
                    if (issuer) bulk_wait_read0();          // previous store has finished reading the unit
                    named_bar_sync(bar_id, 128);
#pragma unroll
                    for (int sub = 0; sub < UCOLS / 32; ++sub) {
                        float v[32];
                        tmem_ld32(taddr + u * UCOLS + sub * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int cj = col + sub * 32 + j * 8;
                            float o[8];
                            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                            if (bias && cj + 8 <= p.n_out) {
                                b0 = __ldg(reinterpret_cast<const float4*>(bias + cj));
                                b1 = __ldg(reinterpret_cast<const float4*>(bias + cj + 4));
                            }
                            o[0] = v[j * 8 + 0] + b0.x; o[1] = v[j * 8 + 1] + b0.y; o[2] = v[j * 8 + 2] + b0.z;
                            o[3] = v[j * 8 + 3] + b0.w; o[4] = v[j * 8 + 4] + b1.x; o[5] = v[j * 8 + 5] + b1.y;
                            o[6] = v[j * 8 + 6] + b1.z; o[7] = v[j * 8 + 7] + b1.w;
                            if constexpr (EPI == DN_EPI_BF16) {
                                if (MODE == 1 && pass) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) o[i] -= round_bf16(o[i]);
                                }
                                const int chunk = sub * 4 + j;  // 16-byte chunk = 8 x 16 bit
                                *reinterpret_cast<uint4*>(srow + ((chunk ^ sw) << 4)) =
                                    make_uint4(pack16<MODE == 2>(o[0], o[1]), pack16<MODE == 2>(o[2], o[3]),
                                               pack16<MODE == 2>(o[4], o[5]), pack16<MODE == 2>(o[6], o[7]));
                            } else {
                                if (EPI == DN_EPI_F32 && pe_row >= 0 && cj + 8 <= p.n_out) {
                                    const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.pe + pe_row + cj));
                                    const float4 e1 = __ldg(reinterpret_cast<const float4*>(p.pe + pe_row + cj + 4));
                                    o[0] += e0.x; o[1] += e0.y; o[2] += e0.z; o[3] += e0.w;
                                    o[4] += e1.x; o[5] += e1.y; o[6] += e1.z; o[7] += e1.w;
                                }
                                const int chunk = j * 2;        // 16-byte chunk = 4 fp32
                                *reinterpret_cast<float4*>(srow + ((chunk ^ sw) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
                                *reinterpret_cast<float4*>(srow + (((chunk + 1) ^ sw) << 4)) =
                                    make_float4(o[4], o[5], o[6], o[7]);
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id, 128);
                    if (issuer) {
                        store_unit(stage_buf, ocol0 + col + pass * p.out_lo_col, EPI == DN_EPI_RESID);
                        bulk_commit();
                    }
                    }
                }
            } else {
                // GEGLU / WN_GATE: two 128-column accumulator halves -> 128 bf16 output columns per tile = 2 units
                const float* gbr = (EPI == DN_EPI_WN_GATE) ? gb_row(p, pb_, c.g) : nullptr;
                const int col = c.n * 128 + half * 64;  // logical output column of this warpgroup's unit
                if (col < p.n_out) {
                    constexpr bool precise = MODE == 1;   // split output: full-precision erf / tanh / exp, then hi | lo
                    for (int pass = 0; pass < (precise ? 2 : 1); ++pass) {
                    if (issuer) bulk_wait_read0();
                    if constexpr (EPI == DN_EPI_GEGLU) {
                        // this warpgroup's 64 x-bias + 64 gate-bias values of the tile: one global load per thread, shared
                        // through smem (the per-column __ldg chain inside the loop was exposed latency)
                        float* sbias = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 512) + half * 128;
                        const int r = threadIdx.x & 127;
                        const int wr = c.g * p.g_bias + c.n * WT + half * 64 + (r & 63) + (r >> 6) * 128;
                        sbias[r] = p.bias ? __ldg(p.bias + wr) : 0.f;
                    }
                    named_bar_sync(bar_id, 128);
#pragma unroll 1
                    for (int sub = 0; sub < 2; ++sub) {
                        float lo[32], hi[32];
                        tmem_ld32(taddr + half * 64 + sub * 32, lo);
                        tmem_ld32(taddr + 128 + half * 64 + sub * 32, hi);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int cj = col + sub * 32 + j * 8;
                            float o[8];
                            float pa[8], pb[8];  // per-column parameters of the two accumulator halves
                            auto ld8 = [&](const float* qq, float (&d)[8], float fill) {
                                if (qq && cj + 8 <= p.n_out) {
                                    const float4 x0 = __ldg(reinterpret_cast<const float4*>(qq));
                                    const float4 x1 = __ldg(reinterpret_cast<const float4*>(qq + 4));
                                    d[0] = x0.x; d[1] = x0.y; d[2] = x0.z; d[3] = x0.w;
                                    d[4] = x1.x; d[5] = x1.y; d[6] = x1.z; d[7] = x1.w;
                                } else {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) d[i] = fill;
                                }
                            };
                            if constexpr (EPI == DN_EPI_GEGLU) {
                                const float* sbias = reinterpret_cast<const float*>(smem + STAGES * STAGE_BYTES + 512) + half * 128;
                                const int cc = sub * 32 + j * 8;
                                const float4 a0 = *reinterpret_cast<const float4*>(sbias + cc), a1 = *reinterpret_cast<const float4*>(sbias + cc + 4);
                                const float4 b0 = *reinterpret_cast<const float4*>(sbias + 64 + cc), b1 = *reinterpret_cast<const float4*>(sbias + 64 + cc + 4);
                                pa[0] = a0.x; pa[1] = a0.y; pa[2] = a0.z; pa[3] = a0.w; pa[4] = a1.x; pa[5] = a1.y; pa[6] = a1.z; pa[7] = a1.w;
                                pb[0] = b0.x; pb[1] = b0.y; pb[2] = b0.z; pb[3] = b0.w; pb[4] = b1.x; pb[5] = b1.y; pb[6] = b1.z; pb[7] = b1.w;
                                if constexpr (precise) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) o[i] = gelu_erf(hi[j * 8 + i] + pb[i]) * (lo[j * 8 + i] + pa[i]);
                                } else {
#pragma unroll
                                    for (int i = 0; i < 8; i += 2)
                                        geglu2(hi[j * 8 + i], hi[j * 8 + i + 1], pb[i], pb[i + 1], lo[j * 8 + i], lo[j * 8 + i + 1],
                                               pa[i], pa[i + 1], o[i], o[i + 1]);
                                }
                            } else {
                                const int oc = c.g * p.g_bias + cj;
                                float ga[8], be[8];
                                ld8(p.bias ? p.bias + oc : nullptr, pa, 0.f);
                                ld8(p.bias2 ? p.bias2 + oc : nullptr, pb, 0.f);
                                ld8(gbr ? gbr + cj : nullptr, ga, 1.f);
                                ld8(gbr ? gbr + p.gb_half + cj : nullptr, be, 0.f);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float uu = fmaf(lo[j * 8 + i] + pa[i], ga[i], be[i]);
                                    const float gt = precise ? tanhf(uu) / (1.f + expf(-uu)) : wn_gate(uu);
                                    o[i] = gt + hi[j * 8 + i] + pb[i];
                                }
                            }
                            if (precise && pass) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) o[i] -= round_bf16(o[i]);
                            }
                            const int chunk = sub * 4 + j;
                            *reinterpret_cast<uint4*>(srow + ((chunk ^ sw) << 4)) =
                                make_uint4(pack16<MODE == 2>(o[0], o[1]), pack16<MODE == 2>(o[2], o[3]),
                                           pack16<MODE == 2>(o[4], o[5]), pack16<MODE == 2>(o[6], o[7]));
                        }
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id, 128);
                    if (issuer) {
                        store_unit(stage_buf, ocol0 + col + pass * p.out_lo_col, false);
                        bulk_commit();
                    }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTAS == 2 && !leader) mbar_arrive_remote(&tempty[as], 0);
                else mbar_arrive(&tempty[as]);
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
        if (issuer) bulk_wait_all0();  // smem must stay valid (and writes complete) until the bulk stores are done
    }

    tc_fence_before();
    if (CTAS == 2) {
        cluster_sync_all();   // the peer's barriers / smem must outlive the leader's last multicast and MMA reads
        if (warp == 2) tmem_dealloc2(tmem_base, 512);
    } else {
        __syncthreads();
        if (warp == 2) tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------------
// SIMT checker: one thread per output element, identical semantics, used only by tests / bring-up.
// ------------------------------------------------------------------------------------------------------
__global__ void gemm_check_kernel(const dn_gemm_desc p) {
    const long long total = (long long)p.groups * p.B * p.T * p.n_out;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int oc = (int)(idx % p.n_out);
        long long r = idx / p.n_out;
        const int t = (int)(r % p.T);
        r /= p.T;
        const int b = (int)(r % p.B);
        const int g = (int)(r / p.B);
        const int d = p.dilation << (p.dilation_shl_group ? g : 0);
        const bool dual = p.epi == DN_EPI_GEGLU || p.epi == DN_EPI_WN_GATE;
        const int wrow_lo = g * p.g_w_row + (dual ? (oc / 128) * WT + (oc % 128) : oc);
        const int wrow_hi = wrow_lo + 128;
        const uint16_t* A = reinterpret_cast<const uint16_t*>(p.A);
        const uint16_t* W = reinterpret_cast<const uint16_t*>(p.W);
        float lo = 0.f, hi = 0.f;
        for (int s = 0; s < p.num_segs; ++s) {
            const dn_gemm_seg sg = p.seg[s];
            const int ts = t - sg.shift_mul * d;
            if (ts < 0) continue;
            const uint16_t* a = A + (long long)b * p.a_batch_stride + (long long)ts * p.lda + sg.a_col0 + g * p.g_a_col;
            const uint16_t* wl = W + (long long)wrow_lo * p.ldw + sg.w_k0;
            const uint16_t* wh = W + (long long)wrow_hi * p.ldw + sg.w_k0;
            const int klen = sg.k_blocks * BK;
            for (int k = 0; k < klen; ++k) {
                // columns past the tensor extent read as zero (TMA out-of-bounds fill)
                const float av = (sg.a_col0 + g * p.g_a_col + k < p.a_cols) ? load16(a[k], p.a_fmt) : 0.f;
                lo += av * load16(wl[k], p.w_fmt);
                if (dual && sg.n_mma == 0) hi += av * load16(wh[k], p.w_fmt);
            }
        }
        const long long o = (long long)b * p.out_batch_stride + (long long)t * p.ldo + g * p.g_out_col + oc;
        // 16-bit outputs: bf16 | fp16 | split bf16 pair (hi at o, lo = v - hi at o + out_lo_col)
        auto store16 = [&](float v) {
            if (p.out_lo_col) {
                const __nv_bfloat16 h = __float2bfloat16(v);
                reinterpret_cast<__nv_bfloat16*>(p.out)[o] = h;
                reinterpret_cast<__nv_bfloat16*>(p.out)[o + p.out_lo_col] = __float2bfloat16(v - __bfloat162float(h));
            } else if (p.out_fmt == DN_FMT_F16) {
                reinterpret_cast<__half*>(p.out)[o] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
            } else {
                reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16(v);
            }
        };
        if (p.epi == DN_EPI_BF16) {
            store16(lo + (p.bias ? p.bias[g * p.g_bias + oc] : 0.f));
        } else if (p.epi == DN_EPI_F32) {
            float v = lo + (p.bias ? p.bias[g * p.g_bias + oc] : 0.f);
            if (p.pe) {
                int pos = t + 1;
                if (p.lengths && t >= p.lengths[b]) pos = 0;
                v += p.pe[(long long)pos * p.n_out + oc];
            }
            reinterpret_cast<float*>(p.out)[o] = v;
        } else if (p.epi == DN_EPI_ARGMAX) {
            // slow but simple: every class writes through a per-row scan done by the thread of class 0 of each 128-column part
            if (oc % 128 == 0 && oc < p.n_classes) {
                float bv = -INFINITY;
                int bi = 0x7fffffff;
                for (int cc = oc; cc < oc + 128 && cc < p.n_classes; ++cc) {
                    float acc = 0.f;
                    for (int s2 = 0; s2 < p.num_segs; ++s2) {
                        const dn_gemm_seg sg = p.seg[s2];
                        const int ts = t - sg.shift_mul * d;
                        if (ts < 0 || ts >= p.T) continue;
                        const uint16_t* a = A + (long long)b * p.a_batch_stride + (long long)ts * p.lda + sg.a_col0;
                        const uint16_t* w = W + (long long)cc * p.ldw + sg.w_k0;
                        for (int k = 0; k < sg.k_blocks * BK; ++k)
                            acc += ((sg.a_col0 + k < p.a_cols) ? load16(a[k], p.a_fmt) : 0.f) * load16(w[k], p.w_fmt);
                    }
                    const float x = acc + (p.bias ? p.bias[cc] : 0.f);
                    const bool xn = x != x, bn = bv != bv;
                    if (bi == 0x7fffffff || (xn && !bn) || (!bn && x > bv)) { bv = x; bi = cc; }
                }
                float* po = reinterpret_cast<float*>(p.out) + ((long long)b * p.T + t) * p.ldo + (oc / 128) * 2;
                po[0] = bv;
                po[1] = __int_as_float(bi);
            }
        } else if (p.epi == DN_EPI_DDIM) {
            const float* cf = p.coef + (long long)(p.t_idx ? p.t_idx[0] : 0) * 8;
            const float eh = lo + (p.bias ? p.bias[oc] : 0.f);
            const float xv = reinterpret_cast<float*>(p.out)[o];
            const float x0 = (xv - cf[1] * eh) / fmaxf(cf[0], 1e-10f);
            const float pn = (xv - cf[0] * x0) / fmaxf(cf[1], 1e-10f);
            const float v = x0 * cf[2] + cf[3] * pn;
            reinterpret_cast<float*>(p.out)[o] = v;
            const long long so = ((long long)b * p.T + t) * p.aux_ld + oc;
            const __nv_bfloat16 h = __float2bfloat16(v);
            reinterpret_cast<__nv_bfloat16*>(p.aux)[so] = h;
            if (p.aux_lo_col) reinterpret_cast<__nv_bfloat16*>(p.aux)[so + p.aux_lo_col] = __float2bfloat16(v - __bfloat162float(h));
        } else if (p.epi == DN_EPI_RESID) {
            reinterpret_cast<float*>(p.out)[o] += lo + (p.bias ? p.bias[g * p.g_bias + oc] : 0.f);
        } else if (p.epi == DN_EPI_GEGLU) {
            const int wr = g * p.g_bias + (oc / 128) * WT + (oc % 128);
            const float x = lo + (p.bias ? p.bias[wr] : 0.f);
            const float gt = hi + (p.bias ? p.bias[wr + 128] : 0.f);
            store16(gelu_erf(gt) * x);
        } else {
            float u = lo + (p.bias ? p.bias[g * p.g_bias + oc] : 0.f);
            const float rr = hi + (p.bias2 ? p.bias2[g * p.g_bias + oc] : 0.f);
            const float* gbr = gb_row(p, b, g);
            if (gbr) u = u * gbr[oc] + gbr[p.gb_half + oc];
            store16(tanhf(u) / (1.f + expf(-u)) + rr);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

static int encode_map(CUtensorMap* m, CUtensorMapDataType dt, const void* base, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box);
int encode_bf16_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    return encode_map(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}
int encode_f32_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box) {
    return encode_map(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}
static int encode_map(CUtensorMap* m, CUtensorMapDataType dt, const void* base, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return DN_EDRIVER;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : DN_EINVAL;
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DN_PDL");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static int g_sm_limit = 0;   // dn_set_sm_limit: SMs the persistent kernels may occupy (0 = all)
int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

template <int EPI, int CTAS, int MODE>
static int launch_tc(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& w128, const CUtensorMap& w64,
                     const CUtensorMap& o, const CUtensorMap& ac, const CUtensorMap& oc, const dn_gemm_desc& d, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        DN_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<EPI, CTAS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
        attr_set = true;
    }
    const long long total = (long long)d.groups * make_tiling(d, BM * CTAS).m_tiles * d.n_tiles;
    const int units = num_sms() / CTAS;   // CTAs (or CTA pairs) resident at once
    const int grid = (int)(total < units ? total : units) * CTAS;
    DN_CUDA_OK(launch_ex(gemm_tc_kernel<EPI, CTAS, MODE>, grid, GEMM_THREADS, GEMM_SMEM, st, CTAS, a, w, w128, w64, o, ac, oc, d));
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}

}  // namespace dn

using namespace dn;

extern "C" int dn_gemm(const dn_gemm_desc* dp, int32_t impl, void* stream) {
    if (!dp || !dp->A || !dp->W || !dp->out) return DN_EINVAL;
    const dn_gemm_desc& d = *dp;
    if (d.B <= 0 || d.T <= 0 || d.groups <= 0 || d.num_segs <= 0 || d.num_segs > DN_MAX_SEGS || d.n_tiles <= 0) return DN_EINVAL;
    if (d.n_out % 8 || d.lda % 8 || d.ldw % 8 || d.w_rows % 16) return DN_EINVAL;
    if ((reinterpret_cast<uintptr_t>(d.A) | reinterpret_cast<uintptr_t>(d.W) | reinterpret_cast<uintptr_t>(d.out)) & 15)
        return DN_EINVAL;
    const bool f32out = d.epi == DN_EPI_F32 || d.epi == DN_EPI_RESID || d.epi == DN_EPI_DDIM || d.epi == DN_EPI_ARGMAX;
    if (d.epi == DN_EPI_ARGMAX && (d.groups != 1 || d.n_classes <= 0 || d.n_classes > d.n_out || d.ldo < 4 * d.n_tiles || d.ldo % 2 ||
                                   d.out_batch_stride != (long long)d.T * d.ldo))
        return DN_EINVAL;
    if (d.ldo % (f32out ? 4 : 8) || d.g_out_col % 8) return DN_EINVAL;
    if (d.epi == DN_EPI_DDIM && (!d.coef || !d.aux || !d.t_idx || d.groups != 1 || d.n_tiles != 1 || d.n_out % 16 || d.n_out > 256 ||
                                 d.aux_ld % 8 || d.aux_lo_col % 8 || d.out_batch_stride != (long long)d.T * d.ldo ||
                                 ((reinterpret_cast<uintptr_t>(d.aux) | reinterpret_cast<uintptr_t>(d.bias)) & 15)))
        return DN_EINVAL;
    if ((unsigned)d.a_fmt > 1u || (unsigned)d.w_fmt > 1u || (unsigned)d.out_fmt > 1u || d.out_lo_col < 0) return DN_EINVAL;
    if (d.out_lo_col && (f32out || d.n_out % 64 || d.out_lo_col % 8 || d.out_fmt != DN_FMT_BF16)) return DN_EINVAL;
    if (d.out_fmt == DN_FMT_F16 && f32out) return DN_EINVAL;
    if (d.row_chunk != 0 && d.row_chunk != 8 && d.row_chunk != 16 && d.row_chunk != 32 && d.row_chunk != 64) return DN_EINVAL;
    if (d.a_fmt != d.w_fmt) return DN_EINVAL;   // kind::f16 MMAs take both operands in ONE 16-bit format (mixing faults)
    const int mode = d.out_lo_col ? 1 : (d.out_fmt == DN_FMT_F16 ? 2 : 0);
    for (int s = 0; s < d.num_segs; ++s)
        if (d.seg[s].k_blocks <= 0 || (d.seg[s].n_mma != 0 && d.seg[s].n_mma != 128) || d.seg[s].a_col0 % 8 ||
            d.seg[s].w_k0 % 8)
            return DN_EINVAL;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    if (impl == DN_GEMM_SIMT_CHECK) {
        const long long total = (long long)d.groups * d.B * d.T * d.n_out;
        int grid = (int)((total + 255) / 256 > 148 * 64 ? 148 * 64 : (total + 255) / 256);
        gemm_check_kernel<<<grid, 256, 0, st>>>(d);
        DN_LAUNCH_CHECK();
        count_launch();
        return 0;
    }

    CUtensorMap ma, mw, mw128, mw64, mac, moc;
    {
        cuuint64_t dims[3] = {(cuuint64_t)d.a_cols, (cuuint64_t)d.T, (cuuint64_t)d.B};
        cuuint64_t str[2] = {(cuuint64_t)d.lda * 2, (cuuint64_t)d.a_batch_stride * 2};
        cuuint32_t box[3] = {BK, BM, 1};
        int r = encode_bf16_map(&ma, d.A, 3, dims, str, box);
        if (r) return r;
        mac = ma;
        if (d.row_chunk) {   // packed rows: one box per row chunk for the tiles that straddle utterances
            box[1] = (cuuint32_t)d.row_chunk;
            r = encode_bf16_map(&mac, d.A, 3, dims, str, box);
            if (r) return r;
        }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)d.ldw, (cuuint64_t)d.w_rows};
        cuuint64_t str[1] = {(cuuint64_t)d.ldw * 2};
        cuuint32_t box[2] = {BK, WT};
        int r = encode_bf16_map(&mw, d.W, 2, dims, str, box);
        if (r) return r;
        cuuint32_t box2[2] = {BK, 128};
        r = encode_bf16_map(&mw128, d.W, 2, dims, str, box2);
        if (r) return r;
        cuuint32_t box3[2] = {BK, 64};
        r = encode_bf16_map(&mw64, d.W, 2, dims, str, box3);
        if (r) return r;
    }
    CUtensorMap mo;
    {
        // output map: columns are clipped at the logical extent so partial tiles never touch neighbouring data
        const int esz = f32out ? 4 : 2;
        const long long ocols = (d.groups > 1 ? (long long)(d.groups - 1) * d.g_out_col + d.n_out : d.n_out) + d.out_lo_col;
        cuuint64_t dims[3] = {(cuuint64_t)ocols, (cuuint64_t)d.T, (cuuint64_t)d.B};
        cuuint64_t str[2] = {(cuuint64_t)d.ldo * esz, (cuuint64_t)d.out_batch_stride * esz};
        cuuint32_t box[3] = {(cuuint32_t)(f32out ? 32 : 64), BM, 1};
        int r = encode_map(&mo, f32out ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d.out, 3,
                           dims, str, box);
        if (r) return r;
        moc = mo;
        if (d.row_chunk) {
            box[1] = (cuuint32_t)d.row_chunk;
            r = encode_map(&moc, f32out ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d.out, 3, dims, str, box);
            if (r) return r;
        }
    }
#define DN_LAUNCH_TC(EPI, CT, MD) launch_tc<EPI, CT, MD>(ma, mw, mw128, mw64, mo, mac, moc, d, st)
#define DN_BY_CTAS(EPI, MD) (impl == DN_GEMM_TCGEN05_2CTA ? DN_LAUNCH_TC(EPI, 2, MD) : DN_LAUNCH_TC(EPI, 1, MD))
    switch (d.epi) {
        case DN_EPI_BF16: return mode == 1 ? DN_BY_CTAS(DN_EPI_BF16, 1) : mode == 2 ? DN_BY_CTAS(DN_EPI_BF16, 2) : DN_BY_CTAS(DN_EPI_BF16, 0);
        case DN_EPI_F32: return DN_BY_CTAS(DN_EPI_F32, 0);
        case DN_EPI_RESID: return DN_BY_CTAS(DN_EPI_RESID, 0);
        case DN_EPI_DDIM: return DN_BY_CTAS(DN_EPI_DDIM, 0);
        case DN_EPI_ARGMAX: return DN_BY_CTAS(DN_EPI_ARGMAX, 0);
        case DN_EPI_GEGLU: return mode == 1 ? DN_BY_CTAS(DN_EPI_GEGLU, 1) : mode == 2 ? DN_BY_CTAS(DN_EPI_GEGLU, 2) : DN_BY_CTAS(DN_EPI_GEGLU, 0);
        case DN_EPI_WN_GATE: return mode == 1 ? DN_BY_CTAS(DN_EPI_WN_GATE, 1) : mode == 2 ? DN_BY_CTAS(DN_EPI_WN_GATE, 2) : DN_BY_CTAS(DN_EPI_WN_GATE, 0);
        default: return DN_EINVAL;
    }
}

// Host-side statement of the kernel's M tiling (the same make_tiling / decode_tile the device code runs): which (utterance,
// frame) each of the 128 accumulator rows of CTA `cta_rank` of M tile `m_tile` holds.  Rows past T or past the batch are
// reported as they are computed (b may equal B, t may be >= T): the TMA loads zero-fill them and the stores are clipped.
extern "C" int dn_gemm_tile_rows(int32_t B, int32_t T, int32_t row_chunk, int32_t ctas, int32_t m_tile, int32_t cta_rank,
                                 int32_t* m_tiles, int32_t* b_out, int32_t* t_out, int32_t* boxes) {
    if (B <= 0 || T <= 0 || (ctas != 1 && ctas != 2) || cta_rank < 0 || cta_rank >= ctas) return DN_EINVAL;
    if (row_chunk != 0 && row_chunk != 8 && row_chunk != 16 && row_chunk != 32 && row_chunk != 64) return DN_EINVAL;
    dn_gemm_desc d{};
    d.B = B; d.T = T; d.row_chunk = row_chunk; d.n_tiles = 1; d.groups = 1;
    const Tiling tg = make_tiling(d, BM * ctas);
    if (m_tiles) *m_tiles = tg.m_tiles;
    if (m_tile < 0 || m_tile >= tg.m_tiles) return (b_out || t_out) ? DN_EINVAL : 0;
    const TileCoord c = decode_tile(d, m_tile, tg, BM * ctas, cta_rank * BM);
    if (boxes) *boxes = c.contig ? 1 : BM / tg.ch;
    for (int row = 0; row < BM; ++row) {
        int wb = c.b, t = c.t0 + row;
        if (!c.contig) {
            const int chn = c.chunk0 + row / tg.ch;
            wb = chn / tg.per_t;
            t = (chn - wb * tg.per_t) * tg.ch + row % tg.ch;
        }
        if (b_out) b_out[row] = wb;
        if (t_out) t_out[row] = t;
    }
    return 0;
}

extern "C" int dn_abi_version(void) { return 2; }
extern "C" int dn_set_sm_limit(int32_t n) {
    if (n < 0 || (n & 1)) return DN_EINVAL;   // even: CTA pairs occupy two SMs
    dn::g_sm_limit = n;
    return 0;
}
extern "C" unsigned long long dn_launch_count(void) { return dn::g_launch_count; }
