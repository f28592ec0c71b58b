// p24_common.cuh — workspace layout, GT record layout and block-level selection primitives shared
// by the kernels of libp24_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#include "../../include/p24.h"
#include "p24_math.cuh"

#define P24_THREADS 256
#define P24_MAX_LEVELS 4   // feature levels of the anchor grid
#define P24_WSIDE 7        // a GT's centre window lies inside a 7 x 7 block of grid cells per level (5 x 5 pass the test)
#define P24_WSLOTS (P24_WSIDE * P24_WSIDE)
#define P24_WT_HDR (P24_WSLOTS * P24_MAX_LEVELS)   // wtab row: [slot costs | ix0, iy0 per level (int bits)]
#define P24_WT_STRIDE (P24_WT_HDR + 2 * P24_MAX_LEVELS + 4)   // 208 floats
#define P24_WARPS (P24_THREADS / 32)
#define P24_LISTCAP 4096   // entries a GT's top-10 list can hold (more -> brute-force path of k_tail)

// ---- per-GT record (floats): k_prep writes it, the seed items of k_pass add [4] and [58] ----------------------
// [0] cx  [1] cy  [2] rin2  [3] rrej2   (one 128-bit shared-memory load for the per-pair tests)
// [4] far2  [5] class (as float)  [6] rgmax  [7] rgmin
// [8..31] vertex x  [32..55] vertex y  [56] mean rg^2  [57] mean rg  [58] T  [59] pad  [60..83] ray length rg
// [84..91] origin (column, row; int bits) of the GT's 7 x 7 centre-window block on every level
#define GT_CX 0
#define GT_CY 1
#define GT_RIN2 2    // squared radius of a disc around (cx, cy) that lies inside the polygon (0: none)
#define GT_RREJ2 3   // squared radius beyond which the angle sum is provably < 349 degrees
#define P24_FQ_PER_GT 256
#define GT_FAR2 4    // squared centre distance below which no prediction's pair value can reach T (0: none)
#define GT_CLS 5
#define GT_RGMAX 6
#define GT_RGMIN 7
#define GT_VX 8
#define GT_VY 32
#define GT_RGMS 56    // mean of rg^2
#define GT_RGMEAN 57  // mean of rg
#define GT_T 58       // certified lower bound of the GT's 10th largest pair value over the candidates (-inf: none)
#define GT_RG 60
#define GT_ORG 84
#define GT_REC 92

static inline size_t p24_align(size_t x) { return (x + 255) & ~(size_t)255; }

// status words (workspace): sticky error bits + path statistics, read back by p24_read_status
#define ST_ERR 0        // P24_ERR_* bits
#define ST_BRUTE 1      // GTs whose dynamic k took the brute-force path (cumulative)
#define ST_SPILL 2      // GTs that spilled into the penalised regime (cumulative)
#define ST_LISTMAX 3    // longest top-10 list seen
#define ST_WAITCYC 4    // clock cycles the last fused all-reduce waited for its peers
#define ST_LISTSUM 5    // list entries seen (cumulative) ...
#define ST_GTS 6        // ... over this many GTs
#define ST_EXACT 7      // GTs whose dynamic k needed exact pair values (the bound bracket straddled an integer)
#define ST_WORDS 8

struct P24Workspace {
    size_t ticket;      // [16] unsigned: tile-queue head of k_pass, largest num_gt, completion count of k_tail, far-queue entries,
                        //               window-queue head, number of centre-window pairs of the batch, far-queue head, tiles done
    size_t acc_fix;     // [28] int64   fixed-point loss sums of the batch (zero between calls)
    size_t status;      // [ST_WORDS] int
    // (what k_prep writes exists twice, by step parity: it runs beside the previous step's k_tail)
    size_t seed_done;   // [2][B] int   seed items of the image that are complete (zero between calls)
    size_t ngt;         // [2][B] int   num_gt of the image
    size_t ncand;       // [B] int      candidate anchors of the image (zero between calls)
    size_t lcount;      // [B, Lmax] int   entries in the GT's top-10 list (zero between calls)
    size_t gt_rec;      // [B, Lmax, GT_REC] float
    size_t wtab;        // [B, Lmax, P24_WT_STRIDE] float  SimOTA cost of the GT's centre-window anchors by window slot
                        //                                  (+inf: not in the window / not in the polygon) + the window origins
    size_t rare;        // [B] int      GTs of the image that need a rare path of k_tail (zero between calls)
    size_t list;        // [B, Lmax, P24_LISTCAP] float2   (value bound, anchor | all-apart flag) of the GT's far candidates whose bound
                        //                                  reaches T (arrival order)
    size_t claimg;      // [B, Lmax * 10] int   anchors claimed by the GTs (-1: unused slot)
    size_t kreq;        // [B, Lmax] int
    size_t ntake;       // [B, Lmax] int
    size_t cbits;       // [B, tiles * 8] unsigned         candidate bitmap (bit l of word w: anchor 32 w + l)
    size_t brute;       // [B, 8, 10] float   per-CTA partial top-10 of the brute-force path of k_tail
    size_t wlist;       // [B * Lmax * 100] int2           the batch's centre-window pairs (GT slot, anchor), written by k_prep
    size_t fq;          // [B * Lmax * P24_FQ_PER_GT] int2   the batch's far queue (GT slot, anchor): far pairs that did not fit a
                        //                                    tile's own work list (k_pass)
    size_t total;
};

static inline int p24_tiles(int A) { return (A + P24_THREADS - 1) / P24_THREADS; }

static inline P24Workspace p24_layout(int B, int A, int Lmax) {
    P24Workspace w;
    size_t off = 0;
    const size_t BL = (size_t)B * (size_t)Lmax;
    const size_t NB = (size_t)B * (size_t)p24_tiles(A);
    // the counters that must be zero between calls come first (p24_workspace_init clears everything)
    w.ticket = off;     off = p24_align(off + 16 * sizeof(unsigned));
    w.acc_fix = off;    off = p24_align(off + 28 * sizeof(long long));
    w.status = off;     off = p24_align(off + ST_WORDS * sizeof(int));
    w.seed_done = off;  off = p24_align(off + 2 * (size_t)B * sizeof(int));
    w.ngt = off;        off = p24_align(off + 2 * (size_t)B * sizeof(int));
    w.ncand = off;      off = p24_align(off + (size_t)B * sizeof(int));
    w.rare = off;       off = p24_align(off + (size_t)B * sizeof(int));
    w.lcount = off;     off = p24_align(off + BL * sizeof(int));
    w.gt_rec = off;     off = p24_align(off + 2 * BL * GT_REC * sizeof(float));
    w.wtab = off;       off = p24_align(off + 2 * BL * P24_WT_STRIDE * sizeof(float));
    w.list = off;       off = p24_align(off + BL * P24_LISTCAP * 2 * sizeof(float));
    w.claimg = off;     off = p24_align(off + BL * P24_TOPK * sizeof(int));
    w.kreq = off;       off = p24_align(off + BL * sizeof(int));
    w.ntake = off;      off = p24_align(off + BL * sizeof(int));
    w.cbits = off;      off = p24_align(off + NB * P24_WARPS * sizeof(unsigned));
    w.brute = off;      off = p24_align(off + (size_t)B * 8 * P24_TOPK * sizeof(float));
    w.wlist = off;      off = p24_align(off + 2 * BL * 25 * P24_MAX_LEVELS * 2 * sizeof(int));
    w.fq = off;         off = p24_align(off + BL * P24_FQ_PER_GT * 2 * sizeof(int));
    w.total = off;
    return w;
}

#ifdef __CUDACC__
// ---- (value, index) selection --------------------------------------------------------------
struct KV {
    float v;
    int i;
};

// larger value wins; ties -> smaller index.  NaN never wins.
__device__ __forceinline__ bool kv_gt(float v1, int i1, float v2, int i2) {
    return (v1 > v2) || (v1 == v2 && i1 < i2);
}
// smaller value wins; ties -> smaller index.
__device__ __forceinline__ bool kv_lt(float v1, int i1, float v2, int i2) {
    return (v1 < v2) || (v1 == v2 && i1 < i2);
}

template <bool MAX>
__device__ __forceinline__ KV warp_select(KV x) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float v2 = __shfl_xor_sync(0xffffffffu, x.v, off);
        const int i2 = __shfl_xor_sync(0xffffffffu, x.i, off);
        const bool take = MAX ? kv_gt(v2, i2, x.v, x.i) : kv_lt(v2, i2, x.v, x.i);
        if (take) {
            x.v = v2;
            x.i = i2;
        }
    }
    return x;
}

// Block-wide arg-select over per-thread candidates; every thread gets the winner.
// s_red: P24_WARPS entries of shared scratch.  Contains two __syncthreads().
template <bool MAX>
__device__ __forceinline__ KV block_select(KV x, KV* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    x = warp_select<MAX>(x);
    __syncthreads();  // protect s_red reuse from a previous call
    if (lane == 0) s_red[warp] = x;
    __syncthreads();
    KV y = s_red[lane < P24_WARPS ? lane : 0];
    y = warp_select<MAX>(y);
    return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ float warp_prod(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// order-preserving map float -> uint32 (smaller float -> smaller key)
__device__ __forceinline__ unsigned p24_ordered(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float p24_unordered(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

#define P24_NEG_INF (-INFINITY)
#define P24_POS_INF (INFINITY)
#endif  // __CUDACC__
