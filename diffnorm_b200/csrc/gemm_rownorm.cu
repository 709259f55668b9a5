// dn_gemm_resid_norm: residual GEMM fused with the NEXT adaptive RMSNorm of the transformer stack (C = 512).
//
//     x[m, :]  += A[m, :] W^T + bias                      (LM:692 / :704, the fp32 residual stream)
//     hb[m, :]  = bf16( x[m, :] / max(||x[m, :]||, 1e-12) * sqrt(512) * gamma_p * gamma_t + beta_t )      (LM:629-639)
//
// The separate norm kernel re-reads the whole residual stream from HBM (2 KB per frame); here the row's sum of squares is
// formed on chip while the residual update is in flight, and the normalised bf16 operand of the next GEMM leaves in the
// same kernel.  A thread-block CLUSTER of two CTAs owns a 128-row tile: CTA r computes columns [256 r, 256 r + 256) into
// one of its two 256-column TMEM accumulators (double-buffered, so the epilogue of tile i overlaps the MMAs of tile i+1),
// and the two CTAs exchange per-row partial sums of squares through distributed shared memory.
//   warp 0   TMA producer: A tile + this CTA's 256-row W tile per 64-wide K block, 3-stage ring
//   warp 1   tcgen05.mma issuer (M 128, N 256)
//   warp 2   TMEM alloc; then issues the hb stores (its own bulk-group stream)
//   warp 3   x stream: TMA-loads 128 x 32 fp32 units of the residual into a 3-slot ring, and stores each slot back once the
//            epilogue has updated it in place
//   warps 4..7   pass 1 of a tile (unit = 32 columns): v = (acc + bias) + x, written in place over x in the ring slot (->
//            TMA store of the new residual) and back into TMEM; row partial sums -> own + peer CTA smem.
//   warps 8..11  pass 2 of the tile before (group = 64 columns): TMEM -> v * inv * gamma + beta -> bf16 staging -> TMA
//            store.  The two passes are separate latency chains (x stream / hb stream), so they run side by side.
// HBM traffic per frame: A row + 2 KB read + 2 KB write of x + 1 KB hb (the un-fused pair reads x a second time).
#include "common.cuh"

namespace dn {

constexpr int RN_BM = 128, RN_BK = 64, RN_C = 512, RN_WT = 256;
constexpr int RN_STAGES = 3;
constexpr int RN_A_BYTES = RN_BM * RN_BK * 2;           // 16 KB
constexpr int RN_B_BYTES = RN_WT * RN_BK * 2;           // 32 KB
constexpr int RN_STAGE_BYTES = RN_A_BYTES + RN_B_BYTES;
constexpr int RN_THREADS = 384;
constexpr int RN_UNIT = 128 * 128;                      // 128 rows x 128 B: 32 fp32 or 64 bf16 columns
constexpr int RN_XSLOTS = 3;
constexpr int RN_UNITS = RN_WT / 32;                    // 8 fp32 units per tile per CTA
constexpr int RN_GROUPS = RN_WT / 64;                   // 4 bf16 groups
// smem map (1024-aligned base): A/W ring | x ring | hb staging | barriers (1 KB) | bias, gamma, beta [256] | ss [4][2][128]
constexpr int RN_OFF_X = RN_STAGES * RN_STAGE_BYTES;
constexpr int RN_OFF_H = RN_OFF_X + RN_XSLOTS * RN_UNIT;
constexpr int RN_OFF_BAR = RN_OFF_H + RN_UNIT;
constexpr int RN_OFF_PAR = RN_OFF_BAR + 1024;
constexpr int RN_SMEM = RN_OFF_PAR + (3 * RN_WT + 8 * RN_BM) * 4 + 1024;
static_assert(RN_SMEM <= 232448, "shared memory budget");

int encode_bf16_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box);
int encode_f32_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box);
int num_sms();

struct RnParams {
    int M, k_blocks;
    const float* bias;
    const float* gamma_p;
    const float* gb;
    long long gb_t_stride;
    const int* t_idx;
};

// store / arrive into the peer CTA's shared memory (same offset), and the cluster-scope acquire that pairs with them
__device__ __forceinline__ void st_cluster_f32(float* local_ptr, uint32_t cta, float v) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.f32 [ra], %2;\n\t"
        "}" ::"r"(smem_u32(local_ptr)),
        "r"(cta), "f"(v)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

#ifdef RN_PROFILE
__device__ unsigned long long rn_prof[16];
#define RN_T0() const long long _t0 = clock64()
#define RN_T1(i) _acc[i] += clock64() - _t0
#define RN_FLUSH(i) atomicAdd(&rn_prof[i], (unsigned long long)_acc[i])
#else
#define RN_T0()
#define RN_T1(i)
#define RN_FLUSH(i)
#endif

__global__ void __launch_bounds__(RN_THREADS, 1)
gemm_resid_norm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH, const RnParams p) {
    extern __shared__ uint8_t rn_smem_raw[];
    // aligned with pointer arithmetic on the __shared__ array (not through an integer), so accesses compile to LDS / STS
    uint8_t* smem = rn_smem_raw + ((1024u - (smem_u32(rn_smem_raw) & 1023u)) & 1023u);
    uint8_t* xring = smem + RN_OFF_X;
    uint8_t* hbuf = smem + RN_OFF_H;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RN_OFF_BAR);
    uint64_t* full = bars;                 // [3] A/W stage landed
    uint64_t* empty = bars + 3;            // [3] A/W stage consumed by the MMAs
    uint64_t* tfull = bars + 6;            // [2] accumulator complete
    uint64_t* tempty = bars + 8;           // [2] accumulator drained (the 4 pass-2 warps)
    uint64_t* xfull = bars + 10;           // [3] x unit landed in ring slot
    uint64_t* xready = bars + 13;          // [3] slot updated in place by the 128 pass-1 threads
    uint64_t* hready = bars + 16;          // hb staging written (128 pass-2 threads)
    uint64_t* hfree = bars + 17;           // hb staging read by its TMA store
    uint64_t* ssfull = bars + 18;          // [4] pass 1 of tile (it & 3) done in BOTH CTAs: 4 local + 4 remote warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
    float* s_bias = reinterpret_cast<float*>(smem + RN_OFF_PAR);
    float* s_gam = s_bias + RN_WT;
    float* s_bet = s_gam + RN_WT;
    float* s_ss = s_bet + RN_WT;           // [tile & 3][source CTA][row]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef RN_PROFILE
    long long _acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long _tk = clock64();
#endif
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int m_tiles = (p.M + RN_BM - 1) / RN_BM;
    const int my_tiles = cluster_id < m_tiles ? (m_tiles - cluster_id + n_clusters - 1) / n_clusters : 0;
    const int col0 = (int)rank * RN_WT;    // this CTA's first output column

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmH);
        for (int i = 0; i < RN_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&ssfull[i], 8);
        for (int i = 0; i < RN_XSLOTS; ++i) {
            mbar_init(&xfull[i], 1);
            mbar_init(&xready[i], 128);
        }
        mbar_init(hready, 128);
        mbar_init(hfree, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    {   // per-column parameters of this CTA's half: bias, gamma_eff = gamma_p * gamma_t, beta_t
        const float* g = p.gb ? p.gb + (long long)p.t_idx[0] * p.gb_t_stride : nullptr;
        for (int c = threadIdx.x; c < RN_WT; c += RN_THREADS) {
            const int gc = col0 + c;
            s_bias[c] = p.bias ? p.bias[gc] : 0.f;
            s_gam[c] = (p.gamma_p ? p.gamma_p[gc] : 1.f) * (g ? g[gc] : 1.f);
            s_bet[c] = g ? g[RN_C + gc] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                    // both CTAs' barriers exist before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int m0 = (cluster_id + it * n_clusters) * RN_BM;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    { RN_T0(); mbar_wait(&empty[stage], phase ^ 1); RN_T1(0); }
                    uint8_t* sa = smem + stage * RN_STAGE_BYTES;
                    mbar_expect_tx(&full[stage], RN_STAGE_BYTES);
                    tma_load_2d(&tmA, &full[stage], sa, kb * RN_BK, m0);
                    tma_load_2d(&tmW, &full[stage], sa + RN_A_BYTES, kb * RN_BK, col0);
                    if (++stage == RN_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t idesc = umma_idesc_bf16_m128(RN_WT);
            for (int it = 0; it < my_tiles; ++it) {
                const int as = it & 1;
                { RN_T0(); mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1); RN_T1(1); }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * RN_WT;
                uint32_t acc = 0;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    { RN_T0(); mbar_wait(&full[stage], phase); RN_T1(2); }
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * RN_STAGE_BYTES);
                    const uint64_t da = umma_desc_sw128(sa);
                    const uint64_t db = umma_desc_sw128(sa + RN_A_BYTES);
#pragma unroll
                    for (int k = 0; k < RN_BK / 16; ++k) {
                        umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, acc);
                        acc = 1;
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == RN_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull[as]);
            }
        }
    } else if (warp == 2) {
        if (lane == 0) {
            // hb stores: groups in order; one staging tile, handed back once the store has read it
            const int total = my_tiles * RN_GROUPS;
            for (int n = 0; n < total; ++n) {
                const int it = n / RN_GROUPS, q = n % RN_GROUPS;
                const int m0 = (cluster_id + it * n_clusters) * RN_BM;
                { RN_T0(); mbar_wait(hready, n & 1); RN_T1(5); }
                tma_store_3d(&tmH, hbuf, col0 + q * 64, m0, 0);
                bulk_commit();
                { RN_T0(); bulk_wait_read0(); RN_T1(6); }
                mbar_arrive(hfree);
            }
            bulk_wait_all0();
        }
    } else if (warp == 3) {
        if (lane == 0) {
            // x stream: unit k = (tile it, unit u) lives in ring slot k % 3
            const int total = my_tiles * RN_UNITS;
            auto load = [&](int k) {
                const int it = k / RN_UNITS, u = k % RN_UNITS, s = k % RN_XSLOTS;
                const int m0 = (cluster_id + it * n_clusters) * RN_BM;
                mbar_expect_tx(&xfull[s], RN_UNIT);
                tma_load_3d(&tmX, &xfull[s], xring + s * RN_UNIT, col0 + u * 32, m0, 0);
                const int kp = k + RN_UNITS;             // one tile ahead into L2: the ring then sees L2 latency, not DRAM's
                if (kp < total)
                    tma_prefetch_l2_3d(&tmX, col0 + (kp % RN_UNITS) * 32, (cluster_id + (kp / RN_UNITS) * n_clusters) * RN_BM, 0);
            };
            for (int k = RN_XSLOTS; k < RN_UNITS && k < total; ++k)
                tma_prefetch_l2_3d(&tmX, col0 + k * 32, cluster_id * RN_BM, 0);
            for (int k = 0; k < RN_XSLOTS && k < total; ++k) load(k);
            for (int k = 0; k < total; ++k) {
                const int it = k / RN_UNITS, u = k % RN_UNITS, s = k % RN_XSLOTS;
                const int m0 = (cluster_id + it * n_clusters) * RN_BM;
                { RN_T0(); mbar_wait(&xready[s], (k / RN_XSLOTS) & 1); RN_T1(3); }
                tma_store_3d(&tmX, xring + s * RN_UNIT, col0 + u * 32, m0, 0);
                bulk_commit();
                { RN_T0(); bulk_wait_read0(); RN_T1(4); }   // slot free again
                if (k + RN_XSLOTS < total) load(k + RN_XSLOTS);
            }
            bulk_wait_all0();
        }
    } else if (warp < 8) {
        // ---- pass 1: residual update in place in the x ring, updated values back into TMEM, partial sums to both CTAs
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int sw = row & 7;
        const uint32_t peer = rank ^ 1u;
        for (int it = 0; it < my_tiles; ++it) {
            const int as = it & 1, par = it & 3;
            const uint32_t taddr = tmem_base + as * RN_WT + ((uint32_t)(q * 32) << 16);
            { RN_T0(); mbar_wait(&tfull[as], (it >> 1) & 1); RN_T1(7); }
            tc_fence_after();
            float ss = 0.f;
#pragma unroll 1
            for (int u = 0; u < RN_UNITS; ++u) {
                const int k = it * RN_UNITS + u;
                const int s = k % RN_XSLOTS;
                float a[32];
                tmem_ld32(taddr + u * 32, a);
                { RN_T0(); mbar_wait(&xfull[s], (k / RN_XSLOTS) & 1); RN_T1(8); }
                uint8_t* xrow = xring + s * RN_UNIT + row * 128;
                { RN_T0(); tmem_ld_wait(); RN_T1(12); }
#ifdef RN_PROFILE
                const long long _tc = clock64();
#endif
                // all loads before any store: the in-place stores may alias the loads as far as the compiler can tell, and a
                // load -> add -> store chain per 16 bytes costs a shared-memory round trip each
                float4 xo[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) xo[c] = *reinterpret_cast<const float4*>(xrow + ((c ^ sw) << 4));
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 bb = *reinterpret_cast<const float4*>(s_bias + u * 32 + c * 4);
                    // (acc + bias) + x: the order of the un-fused path (bias in the epilogue, then the TMA reduce-add)
                    const float v0 = (a[c * 4] + bb.x) + xo[c].x, v1 = (a[c * 4 + 1] + bb.y) + xo[c].y;
                    const float v2 = (a[c * 4 + 2] + bb.z) + xo[c].z, v3 = (a[c * 4 + 3] + bb.w) + xo[c].w;
                    ss += v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3;
                    a[c * 4] = v0, a[c * 4 + 1] = v1, a[c * 4 + 2] = v2, a[c * 4 + 3] = v3;
                }
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4*>(xrow + ((c ^ sw) << 4)) = make_float4(a[c * 4], a[c * 4 + 1], a[c * 4 + 2], a[c * 4 + 3]);
#ifdef RN_PROFILE
                _acc[13] += clock64() - _tc;
#endif
                { RN_T0(); fence_proxy_async_smem(); RN_T1(14); }
                mbar_arrive(&xready[s]);
                { RN_T0(); tmem_st32(taddr + u * 32, a); tmem_st_wait(); RN_T1(15); }
            }
            tmem_st_wait();
            tc_fence_before();                           // the pass-2 warps read these TMEM values after the barrier
            float* slot = s_ss + (par * 2 + (int)rank) * RN_BM + row;
            *slot = ss;
            st_cluster_f32(slot, peer, ss);
            fence_acq_rel_cluster();                     // each lane's stores ordered before lane 0's release below
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_remote_release(&ssfull[par], peer);
                mbar_arrive_remote_release(&ssfull[par], rank);
            }
        }
    } else {
        // ---- pass 2: normalised bf16 rows, 64 columns at a time through the one staging tile
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int sw = row & 7;
        const float scale = sqrtf((float)RN_C);
        for (int it = 0; it < my_tiles; ++it) {
            const int as = it & 1, par = it & 3;
            const uint32_t taddr = tmem_base + as * RN_WT + ((uint32_t)(q * 32) << 16);
            { RN_T0(); mbar_wait_cluster(&ssfull[par], (it >> 2) & 1); RN_T1(9); }
            tc_fence_after();
            const float* sp = s_ss + par * 2 * RN_BM + row;
            const float tot = sp[0] + sp[RN_BM];         // same order in both CTAs
            const float inv = scale / fmaxf(sqrtf(tot), 1e-12f);
#pragma unroll 1
            for (int grp = 0; grp < RN_GROUPS; ++grp) {
                const int n = it * RN_GROUPS + grp;
                uint32_t hv[32];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float a[32];
                    tmem_ld32(taddr + grp * 64 + half * 32, a);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const int cc = grp * 64 + half * 32 + c * 2;
                        const float2 ga = *reinterpret_cast<const float2*>(s_gam + cc);
                        const float2 be = *reinterpret_cast<const float2*>(s_bet + cc);
                        hv[half * 16 + c] = pack_bf16(fmaf(a[c * 2] * inv, ga.x, be.x), fmaf(a[c * 2 + 1] * inv, ga.y, be.y));
                    }
                }
                if (grp == RN_GROUPS - 1) {              // accumulator fully consumed: the MMAs of tile it + 2 may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                }
                { RN_T0(); mbar_wait(hfree, (n & 1) ^ 1); RN_T1(10); }   // the previous group's store has read the staging tile
                uint8_t* hrow = hbuf + row * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4*>(hrow + ((c ^ sw) << 4)) =
                        make_uint4(hv[c * 4], hv[c * 4 + 1], hv[c * 4 + 2], hv[c * 4 + 3]);
                fence_proxy_async_smem();
                mbar_arrive(hready);
            }
        }
    }
#ifdef RN_PROFILE
    if (lane == 0) {
        if (warp == 0) { RN_FLUSH(0); atomicAdd(&rn_prof[11], (unsigned long long)(clock64() - _tk)); }
        if (warp == 1) { RN_FLUSH(1); RN_FLUSH(2); }
        if (warp == 3) { RN_FLUSH(3); RN_FLUSH(4); }
        if (warp == 2) { RN_FLUSH(5); RN_FLUSH(6); }
        if (warp == 4) { RN_FLUSH(7); RN_FLUSH(8); RN_FLUSH(12); RN_FLUSH(13); RN_FLUSH(14); RN_FLUSH(15); }
        if (warp == 8) { RN_FLUSH(9); RN_FLUSH(10); }
    }
#endif
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                    // the peer may still be writing into this CTA's shared memory
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace dn

using namespace dn;

#ifdef RN_PROFILE
// debugging build only: summed wait cycles per role (see RN_T1 indices), zeroed on read
extern "C" int dn_debug_rownorm_profile(unsigned long long* out16) {
    unsigned long long z[16] = {0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, rn_prof, sizeof(z));
    cudaMemcpyToSymbol(rn_prof, z, sizeof(z));
    return 0;
}
#endif

extern "C" int dn_gemm_resid_norm(const dn_resid_norm_desc* dp, void* stream) {
    if (!dp || !dp->A || !dp->W || !dp->x || !dp->hb) return DN_EINVAL;
    const dn_resid_norm_desc& d = *dp;
    if (d.M <= 0 || d.k_blocks <= 0 || d.lda % 8 || d.ldw % 8 || d.lda < d.k_blocks * RN_BK || d.ldw < d.k_blocks * RN_BK)
        return DN_EINVAL;
    if ((reinterpret_cast<uintptr_t>(d.A) | reinterpret_cast<uintptr_t>(d.W) | reinterpret_cast<uintptr_t>(d.x) |
         reinterpret_cast<uintptr_t>(d.hb)) & 15)
        return DN_EINVAL;
    if (d.gb && !d.t_idx) return DN_EINVAL;
    CUtensorMap ma, mw, mx, mh;
    {
        cuuint64_t dims[2] = {(cuuint64_t)d.lda, (cuuint64_t)d.M};
        cuuint64_t str[1] = {(cuuint64_t)d.lda * 2};
        cuuint32_t box[2] = {RN_BK, RN_BM};
        int r = encode_bf16_map(&ma, d.A, 2, dims, str, box);
        if (r) return r;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)d.ldw, (cuuint64_t)RN_C};
        cuuint64_t str[1] = {(cuuint64_t)d.ldw * 2};
        cuuint32_t box[2] = {RN_BK, RN_WT};
        int r = encode_bf16_map(&mw, d.W, 2, dims, str, box);
        if (r) return r;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)RN_C, (cuuint64_t)d.M, 1};
        cuuint64_t str[2] = {(cuuint64_t)RN_C * 4, (cuuint64_t)RN_C * 4 * (cuuint64_t)d.M};
        cuuint32_t box[3] = {32, RN_BM, 1};
        int r = encode_f32_map(&mx, d.x, 3, dims, str, box);
        if (r) return r;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)RN_C, (cuuint64_t)d.M, 1};
        cuuint64_t str[2] = {(cuuint64_t)RN_C * 2, (cuuint64_t)RN_C * 2 * (cuuint64_t)d.M};
        cuuint32_t box[3] = {64, RN_BM, 1};
        int r = encode_bf16_map(&mh, d.hb, 3, dims, str, box);
        if (r) return r;
    }
    static bool attr_set = false;
    if (!attr_set) {
        DN_CUDA_OK(cudaFuncSetAttribute(gemm_resid_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RN_SMEM));
        attr_set = true;
    }
    RnParams p;
    p.M = d.M;
    p.k_blocks = d.k_blocks;
    p.bias = d.bias;
    p.gamma_p = d.gamma_p;
    p.gb = d.gb;
    p.gb_t_stride = d.gb_t_stride;
    p.t_idx = d.t_idx;
    const int m_tiles = (d.M + RN_BM - 1) / RN_BM;
    const int pairs = num_sms() / 2;
    const int clusters = m_tiles < pairs ? m_tiles : pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(RN_THREADS);
    cfg.dynamicSmemBytes = RN_SMEM;
    cfg.stream = reinterpret_cast<cudaStream_t>(stream);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    DN_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_resid_norm_kernel, ma, mw, mx, mh, p));
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}
