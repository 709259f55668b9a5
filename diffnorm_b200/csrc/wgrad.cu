// dn_wgrad: weight-gradient GEMM of the denoiser training step (LM:1514-1613 backward) on tcgen05 tensor cores.
//
//   dW[n, c] += sum_{b, t} dY[b, t, n] * X[b, t - shift, c]          (frames are the contraction dimension)
//
// Both operands are read straight from the row-major [B, T, C] activation layout: the non-contracted dimension
// (channels) is the contiguous one, so both UMMA operands are MN-major.  A TMA box {64 channels, 64 frames, 1 utt}
// lands in smem as 64 rows of 128 B (128-byte swizzle) = one column of 8 "64 x 8" MN-major atoms (SBO = 1024 B
// between 8-frame groups); the next 64 channels are the next box, LBO = 8192 B further on.  A conv tap's gradient is
// the same contraction with X read `shift` frames earlier; the per-utterance 3-D tensor map zero-fills t - shift < 0
// and t >= T, which is exactly the causal padding and keeps utterances apart.
// Groups: `groups` independent problems in one launch (the 8 chains of a WaveNet level): group g reads dY / X at
// column offsets g * g_dy_col / g * g_x_col, optionally with shift << g (chain g has dilation 2^g), and owns dW + g *
// g_dw_stride.
// Split-K: a work item = (128 x 256 output tile, range of 64-frame blocks); partial tiles are added into the fp32
// gradient with TMA reduce-add (the same epilogue as the residual GEMM), so dW must be zeroed (or hold the running
// accumulation) before the launch.
//   warp 0 TMA producer | warp 1 MMA issuer | warp 2 TMEM allocator | warps 4..11 epilogue (2 accumulator stages)
#include "common.cuh"

namespace dn {

constexpr int WG_BM = 128;      // dY channels per tile (UMMA M)
constexpr int WG_BN = 256;      // X channels per tile (UMMA N)
constexpr int WG_BK = 64;       // frames per stage
constexpr int WG_STAGES = 4;
constexpr int WG_BOX = WG_BK * 128;                 // one {64 ch, 64 frames} box: 8 KB
constexpr int WG_A_BYTES = (WG_BM / 64) * WG_BOX;   // 16 KB
constexpr int WG_B_BYTES = (WG_BN / 64) * WG_BOX;   // 32 KB
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_THREADS = 384;
constexpr int WG_EPI_WARPS = 8;
constexpr int WG_UNIT = 128 * 128;                  // epilogue staging unit: 128 rows x 32 fp32
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 1024 + 2 * WG_UNIT + 1024;

int encode_bf16_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box);
int encode_f32_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box);
int num_sms();

// MN-major, 128-byte-swizzled operand: 8-frame groups 1024 B apart (SBO), 64-channel atoms `lbo` bytes apart (LBO)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// bf16 x bf16 -> fp32, M = 128, both operands MN-major
__device__ __forceinline__ uint32_t wg_idesc(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

struct WgItem {
    int m0, n0, blk0, blk1, g;
};

struct WgParams {
    int B, T, shift;
    int dy_col0, x_col0;
    int n_rows, k_cols;
    int m_tiles, n_tiles, splits, blocks_per_utt, total_blocks;
    int groups, g_dy_col, g_x_col, shift_shl_group;
};

__device__ __forceinline__ WgItem wg_decode(const WgParams& p, int item) {
    WgItem w;
    const int per_group = p.m_tiles * p.n_tiles * p.splits;
    w.g = item / per_group;
    item -= w.g * per_group;
    const int tile = item / p.splits, s = item % p.splits;
    w.m0 = (tile / p.n_tiles) * WG_BM;
    w.n0 = (tile % p.n_tiles) * WG_BN;
    const int per = (p.total_blocks + p.splits - 1) / p.splits;
    w.blk0 = s * per;
    w.blk1 = min(p.total_blocks, w.blk0 + per);
    return w;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ CUtensorMap tmOut, const WgParams p) {
    extern __shared__ uint8_t wg_smem_raw[];
    uint8_t* smem = wg_smem_raw + ((1024u - (smem_u32(wg_smem_raw) & 1023u)) & 1023u);   // keeps the shared address space: LDS / STS
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + WG_STAGES;
    uint64_t* tfull = bars + 2 * WG_STAGES;
    uint64_t* tempty = bars + 2 * WG_STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = p.m_tiles * p.n_tiles * p.splits * p.groups;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmY);
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmOut);
        for (int i = 0; i < WG_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], WG_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < total; item += gridDim.x) {
                const WgItem w = wg_decode(p, item);
                const int shift = p.shift << (p.shift_shl_group ? w.g : 0);
                const int ycol = p.dy_col0 + w.g * p.g_dy_col + w.m0, xcol = p.x_col0 + w.g * p.g_x_col + w.n0;
                for (int blk = w.blk0; blk < w.blk1; ++blk) {
                    const int b = blk / p.blocks_per_utt;
                    const int t0 = (blk % p.blocks_per_utt) * WG_BK;
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * WG_STAGE_BYTES;
                    uint8_t* sb = sa + WG_A_BYTES;
                    mbar_expect_tx(&full[stage], WG_STAGE_BYTES);
#pragma unroll
                    for (int i = 0; i < WG_BM / 64; ++i)
                        tma_load_3d(&tmY, &full[stage], sa + i * WG_BOX, ycol + i * 64, t0, b);
#pragma unroll
                    for (int i = 0; i < WG_BN / 64; ++i)
                        tma_load_3d(&tmX, &full[stage], sb + i * WG_BOX, xcol + i * 64, t0 - shift, b);
                    if (++stage == WG_STAGES) {
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
            int as = 0;
            uint32_t aphase = 0;
            for (int item = blockIdx.x; item < total; item += gridDim.x) {
                const WgItem w = wg_decode(p, item);
                int n_mma = p.k_cols - w.n0;
                n_mma = n_mma > WG_BN ? WG_BN : ((n_mma + 15) & ~15);
                const uint32_t idesc = wg_idesc((uint32_t)n_mma);
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * WG_BN;
                uint32_t acc = 0;
                for (int blk = w.blk0; blk < w.blk1; ++blk) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * WG_STAGE_BYTES);
                    const uint64_t da = umma_desc_mn_sw128(sa, WG_BOX);
                    const uint64_t db = umma_desc_mn_sw128(sa + WG_A_BYTES, WG_BOX);
#pragma unroll
                    for (int k = 0; k < WG_BK / 16; ++k) {
                        // 16 frames = two 8-frame groups = 2048 B inside every box
                        umma_bf16(d_tmem, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, acc);
                        acc = 1;
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == WG_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull[as]);
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // epilogue: fp32 accumulator rows -> 128B-swizzled 16 KB units -> TMA reduce-add into dW (clipped at the edges)
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const bool issuer = (warp == 4 + 4 * half) && lane == 0;
        uint8_t* stage_buf = smem + WG_STAGES * WG_STAGE_BYTES + 1024 + half * WG_UNIT;
        uint8_t* srow = stage_buf + row * 128;
        const int sw = row & 7;
        const int bar_id = 1 + half;
        int as = 0;
        uint32_t aphase = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            const WgItem w = wg_decode(p, item);
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + as * WG_BN + ((uint32_t)(q * 32) << 16);
            if (w.blk1 > w.blk0) {
                for (int u = half; u < WG_BN / 32; u += 2) {
                    const int col = w.n0 + u * 32;
                    if (col >= p.k_cols) break;
                    if (issuer) bulk_wait_read0();
                    named_bar_sync(bar_id, 128);
                    float v[32];
                    tmem_ld32(taddr + u * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(srow + ((j ^ sw) << 4)) =
                            make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id, 128);
                    if (issuer) {
                        tma_reduce_add_3d(&tmOut, stage_buf, col, w.m0, w.g);
                        bulk_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
        if (issuer) bulk_wait_all0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace dn

using namespace dn;

extern "C" int dn_wgrad(const dn_wgrad_desc* dp, void* stream) {
    if (!dp || !dp->dY || !dp->X || !dp->dW) return DN_EINVAL;
    const dn_wgrad_desc& d = *dp;
    if (d.B <= 0 || d.T <= 0 || d.n_rows <= 0 || d.k_cols <= 0) return DN_EINVAL;
    if (d.ldy % 8 || d.ldx % 8 || d.ldw % 4 || d.dy_col0 % 8 || d.x_col0 % 8) return DN_EINVAL;
    const int groups = d.groups > 0 ? d.groups : 1;
    if (groups > 1 && (d.g_dy_col % 8 || d.g_x_col % 8 || d.g_dw_stride % 4 || d.g_dw_stride < (int64_t)d.n_rows * d.ldw))
        return DN_EINVAL;
    if ((reinterpret_cast<uintptr_t>(d.dY) | reinterpret_cast<uintptr_t>(d.X) | reinterpret_cast<uintptr_t>(d.dW)) & 15)
        return DN_EINVAL;
    CUtensorMap my, mx, mo;
    {
        cuuint64_t dims[3] = {(cuuint64_t)d.ldy, (cuuint64_t)d.T, (cuuint64_t)d.B};
        cuuint64_t str[2] = {(cuuint64_t)d.ldy * 2, (cuuint64_t)d.dy_batch_stride * 2};
        cuuint32_t box[3] = {64, WG_BK, 1};
        int r = encode_bf16_map(&my, d.dY, 3, dims, str, box);
        if (r) return r;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)d.ldx, (cuuint64_t)d.T, (cuuint64_t)d.B};
        cuuint64_t str[2] = {(cuuint64_t)d.ldx * 2, (cuuint64_t)d.x_batch_stride * 2};
        cuuint32_t box[3] = {64, WG_BK, 1};
        int r = encode_bf16_map(&mx, d.X, 3, dims, str, box);
        if (r) return r;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)d.k_cols, (cuuint64_t)d.n_rows, (cuuint64_t)groups};
        cuuint64_t str[2] = {(cuuint64_t)d.ldw * 4,
                             groups > 1 ? (cuuint64_t)d.g_dw_stride * 4 : (cuuint64_t)d.ldw * 4 * (cuuint64_t)d.n_rows};
        cuuint32_t box[3] = {32, 128, 1};
        int r = encode_f32_map(&mo, d.dW, 3, dims, str, box);
        if (r) return r;
    }
    WgParams p;
    p.B = d.B;
    p.T = d.T;
    p.shift = d.x_shift;
    p.dy_col0 = d.dy_col0;
    p.x_col0 = d.x_col0;
    p.n_rows = d.n_rows;
    p.k_cols = d.k_cols;
    p.m_tiles = (d.n_rows + WG_BM - 1) / WG_BM;
    p.n_tiles = (d.k_cols + WG_BN - 1) / WG_BN;
    p.blocks_per_utt = (d.T + WG_BK - 1) / WG_BK;
    p.total_blocks = d.B * p.blocks_per_utt;
    p.groups = groups;
    p.g_dy_col = d.g_dy_col;
    p.g_x_col = d.g_x_col;
    p.shift_shl_group = d.shift_shl_group;
    const int tiles = p.m_tiles * p.n_tiles * groups;
    int splits = d.splits;
    if (splits <= 0) {  // fill the machine, but keep at least 8 frame blocks per work item
        splits = (2 * num_sms() + tiles - 1) / tiles;
        const int cap = (p.total_blocks + 7) / 8;
        if (splits > cap) splits = cap;
    }
    if (splits < 1) splits = 1;
    if (splits > p.total_blocks) splits = p.total_blocks;
    p.splits = splits;
    static bool attr_set = false;
    if (!attr_set) {
        DN_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
        attr_set = true;
    }
    const long long total = (long long)tiles * splits;  // tiles already counts the groups
    const int grid = (int)(total < num_sms() ? total : num_sms());
    wgrad_tc_kernel<<<grid, WG_THREADS, WG_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(my, mx, mo, p);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}
