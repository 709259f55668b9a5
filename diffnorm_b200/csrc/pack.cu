// dn_pack_weights: the training step's per-step fp32 -> bf16 weight re-packing as ONE launch over a descriptor table
// (include/diffnorm_b200.h, dn_pack_op).  HBM-bound: reads every master weight once per packing that uses it, writes the bf16
// tiles; transposed packings (data-gradient GEMMs) go through a 64 x 65 shared-memory tile so that both the fp32
// reads and the bf16 writes are coalesced.
#include "common.cuh"

namespace dn {

// dst index = row_off(r) + col_off(c) + tap_pos[k] * tap_cols: the block maps are divisions, so they are evaluated once per
// row / once per thread, not per element (the first version spent its time in them: 1.5 TB/s)
__device__ __forceinline__ long long pack_row_off(const dn_pack_op& o, int r) {
    return (o.row0 + (long long)(r / o.rblk) * o.rblk_stride + r % o.rblk) * o.ldd;
}
__device__ __forceinline__ long long pack_col_off(const dn_pack_op& o, int c) {
    return o.col0 + (long long)(c / o.cblk) * o.cblk_stride + c % o.cblk;
}
__device__ __forceinline__ void pack_store(const dn_pack_op& o, long long idx, float v) {
    if (o.out_f32) reinterpret_cast<float*>(o.dst)[idx] = v;
    else reinterpret_cast<__nv_bfloat16*>(o.dst)[idx] = __float2bfloat16_rn(v);
}

constexpr int PK_TILE = 64;   // dst tile: 64 rows x 64 columns (x taps); 256 threads

// two adjacent dst columns per thread: one 4-byte (bf16x2) or 8-byte (fp32x2) store, so a warp writes whole 128-byte lines
__device__ __forceinline__ void pack_store2(const dn_pack_op& o, long long idx, float a, float b) {
    if (o.out_f32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(o.dst) + idx) = make_float2(a, b);
    else *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(o.dst) + idx) = __floats2bfloat162_rn(a, b);
}

__global__ void __launch_bounds__(256) pack_weights_kernel(const dn_pack_op* __restrict__ ops, int n_ops) {
    __shared__ dn_pack_op so;
    __shared__ float tile[PK_TILE][PK_TILE + 1];
    const int bid = blockIdx.x;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_ops - 1;              // last op whose first tile is <= bid
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (ops[mid].tile0 <= bid) lo = mid; else hi = mid - 1;
        }
        so = ops[lo];
    }
    __syncthreads();
    const dn_pack_op& o = so;
    const int t = bid - o.tile0;
    const int r0 = (t / o.tiles_c) * PK_TILE, c0 = (t % o.tiles_c) * PK_TILE;
    // column pairs (c, c + 1) land side by side in dst when every block / offset of the column map is even
    const bool pairs = !((o.cblk | o.cblk_stride | o.col0 | o.ldd | o.tap_cols) & 1) &&
                       !(reinterpret_cast<uintptr_t>(o.dst) & 7);
    if (pairs) {
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 column pairs x 8 rows per pass
        const int c = c0 + 2 * tx;
        const bool two = c + 1 < o.cols;
        const long long coff = c < o.cols ? pack_col_off(o, c) : 0;
        if (!o.src_r_fastest) {
            if (c < o.cols) {
                // a linear weight's row is contiguous along c: 8-byte loads when the element index is even
                const bool vec = two && o.taps == 1 && o.s_col == 1 && !(o.s_row & 1) && !(reinterpret_cast<uintptr_t>(o.src) & 7);
#pragma unroll 4
                for (int i = 0; i < PK_TILE / 8; ++i) {
                    const int r = r0 + ty + 8 * i;
                    if (r < o.rows) {
                        const float* s = o.src + r * o.s_row + c * o.s_col;
                        const long long base = pack_row_off(o, r) + coff;
                        if (vec) {
                            const float2 v = *reinterpret_cast<const float2*>(s);
                            pack_store2(o, base + (long long)o.tap_pos[0] * o.tap_cols, v.x, v.y);
                        } else {
                            for (int k = 0; k < o.taps; ++k) {
                                const long long idx = base + (long long)o.tap_pos[k] * o.tap_cols;
                                if (two) pack_store2(o, idx, s[k * o.s_tap], s[o.s_col + k * o.s_tap]);
                                else pack_store(o, idx, s[k * o.s_tap]);
                            }
                        }
                    }
                }
            }
        } else {
            // src is contiguous along r: one tap at a time through the 64 x 65 tile — reads run along r (64 lanes of a row of
            // the tile = 256 contiguous bytes of src), writes along c in pairs
            const int lx = threadIdx.x & 63, ly = threadIdx.x >> 6;
            const int r = r0 + lx;
            for (int k = 0; k < o.taps; ++k) {
                if (k) __syncthreads();
                if (r < o.rows) {
#pragma unroll 4
                    for (int i = 0; i < PK_TILE / 4; ++i) {
                        const int cc = c0 + ly + 4 * i;
                        if (cc < o.cols) tile[ly + 4 * i][lx] = o.src[r * o.s_row + cc * o.s_col + k * o.s_tap];
                    }
                }
                __syncthreads();
                if (c < o.cols) {
#pragma unroll 4
                    for (int i = 0; i < PK_TILE / 8; ++i) {
                        const int rl = ty + 8 * i, rr = r0 + rl;
                        if (rr < o.rows) {
                            const long long idx = pack_row_off(o, rr) + coff + (long long)o.tap_pos[k] * o.tap_cols;
                            if (two) pack_store2(o, idx, tile[2 * tx][rl], tile[2 * tx + 1][rl]);
                            else pack_store(o, idx, tile[2 * tx][rl]);
                        }
                    }
                }
            }
        }
        return;
    }
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    if (!o.src_r_fastest) {
        const int c = c0 + tx;
        if (c < o.cols) {
            const long long coff = pack_col_off(o, c);
#pragma unroll 4
            for (int i = 0; i < PK_TILE / 4; ++i) {
                const int r = r0 + ty + 4 * i;
                if (r < o.rows) {
                    const float* s = o.src + r * o.s_row + c * o.s_col;
                    const long long base = pack_row_off(o, r) + coff;
                    for (int k = 0; k < o.taps; ++k) pack_store(o, base + (long long)o.tap_pos[k] * o.tap_cols, s[k * o.s_tap]);
                }
            }
        }
    } else {
        // src is contiguous along r: one tap at a time through a 64 x 65 tile (the taps of an element are adjacent in src, so
        // the second and third pass hit the lines the first one brought in)
        const int r = r0 + tx, cw = c0 + tx;
        const long long coff = cw < o.cols ? pack_col_off(o, cw) : 0;
        for (int k = 0; k < o.taps; ++k) {
            if (k) __syncthreads();
            if (r < o.rows) {
#pragma unroll 4
                for (int i = 0; i < PK_TILE / 4; ++i) {
                    const int c = c0 + ty + 4 * i;
                    if (c < o.cols) tile[ty + 4 * i][tx] = o.src[r * o.s_row + c * o.s_col + k * o.s_tap];
                }
            }
            __syncthreads();
            if (cw < o.cols) {
#pragma unroll 4
                for (int i = 0; i < PK_TILE / 4; ++i) {
                    const int rr = r0 + ty + 4 * i;
                    if (rr < o.rows)
                        pack_store(o, pack_row_off(o, rr) + coff + (long long)o.tap_pos[k] * o.tap_cols, tile[tx][ty + 4 * i]);
                }
            }
        }
    }
}

}  // namespace dn

using namespace dn;

extern "C" int dn_pack_weights(const dn_pack_op* ops_device, int32_t n_ops, int32_t total_tiles, void* stream) {
    if (!ops_device || n_ops <= 0 || total_tiles <= 0) return DN_EINVAL;
    pack_weights_kernel<<<total_tiles, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ops_device, n_ops);
    DN_LAUNCH_CHECK();
    count_launch();
    return 0;
}
