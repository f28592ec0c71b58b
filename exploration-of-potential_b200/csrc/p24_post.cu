// p24_post.cu — inference postprocess of YOLOX-24p for sm_100a.
//
// Replaces utils.boxes.postprocess (utils/boxes.py:29-99; twin show_24p.py:212-264), image by image (the reference's
// batched call raises for B >= 2, boxes.py:64-65; the contract is "the per-image result for each i"):
//   class_conf, class_pred = max over the class scores (first maximum on ties)       boxes.py:51
//   keep anchors with obj * class_conf >= conf_thre                                  boxes.py:55
//   rectangle of the 24 decoded points r_k * (theta_k cos theta_k, theta_k sin theta_k) + centre   boxes.py:67-76
//   torchvision nms / batched_nms (coordinate trick) on the rectangles, scores obj * class_conf    boxes.py:78-90
//   rows [cx, cy, r0..r23, obj, class_conf, class_pred] of the kept anchors in NMS (score) order   boxes.py:92
//
// Kernels:
//   k_post_filter  one CTA per 256-anchor tile.  The ONE pass over the prediction: each warp stages its 32 rows
//                  (32 x (27+nc) contiguous floats) in shared memory with a single TMA bulk copy (cp.async.bulk +
//                  mbarrier) when the layout allows it, then every lane reduces its own row (class max / argmax,
//                  score, rectangle) and the tile's candidates are compacted in anchor order.  HBM bound.
//   k_post_nms     one CTA per image: stable descending sort of the candidates by score (bitonic sort of 64-bit
//                  (score, position) keys in shared memory), greedy NMS in chunks of 64 sorted boxes, output rows.
//
// torchvision is third-party arithmetic that the reference does not vendor.  The IoU test repeats
// torchvision's expression  inter / (area_a + area_b - inter) > thr  in fp32 without FMA contraction (this file is
// compiled with -fmad=false), the sort is stable, and batched NMS adds  class * (max_coordinate + 1)  to the boxes in
// fp32 exactly like _batched_nms_coordinate_trick.
#include <stdio.h>
#include <string.h>

#include "p24_common.cuh"
#include "p24_host.h"

namespace {

#define POST_THREADS 256
#define POST_WARPS 8
#define NMS_THREADS 1024
#define SORT_SMEM_MAX 16384  // candidates of one image sorted in shared memory (128 KB of keys)
#define NCELL_MAX 2048
#define SPLIT_GROUPS 128               // most groups of an image that the split sweep handles (one warp per group)
#define META_WORDS (SPLIT_GROUPS + 4)
#define OUT_SPLIT 4                    // CTAs per image of k_post_out

struct PostParams {
    const float* pred;
    long long img_stride, row_stride;
    int B, A, nc, C;
    float coef_x[P24_RAYS], coef_y[P24_RAYS];
    float conf_thre, nms_thre;
    int class_agnostic;
    int32_t* cand_count;
    int32_t* det_count;
    float* det_rows;
    int32_t* keep_idx;
    float* rect_debug;
    // workspace
    int* tcount;     // [B, tiles]
    float* c_score;  // [B, tiles*256]
    float* c_conf;   // [B, tiles*256]
    int* c_anchor;   // [B, tiles*256]
    int* c_cls;      // [B, tiles*256]
    float4* c_rect;  // [B, tiles*256]
    float4* s_rect;  // [B, A]   sorted (and class-shifted) rectangles
    unsigned long long* g_keys;  // [B, npad]  the sorted keys (slot, kept / suppressed bits): sort scratch when an image has more than
                                 //            SORT_SMEM_MAX candidates, and what k_post_nms hands to k_post_sweep / k_post_out
    int* g_meta;     // [B, META_WORDS]  k_post_nms -> k_post_sweep: [0] sweeps pending, [1] groups, [2 ..] group starts
    int* s_order;    // [B, A]   sorted boxes grouped by x cell
    int* s_cell;     // [B, A]   x cell of every sorted box
    float4* c_srect; // [B, A]   the sorted rectangles again, in cell order (coalesced reads of a cell)
    int tiles;
    int npad_global;
    // raw head outputs (pred == NULL): per level reg [B,26,H,W], obj [B,1,H,W], cls [B,nc,H,W] before the sigmoid / decode of
    // YOLOXHead.forward(train=False) + decode_outputs (models/yolo_head_24p.py:191, 201-211, 239-256)
    const float* raw[3][4];
    long long raw_bs[3][4];
    int nlev;
    int lev_off[5], lev_w[4];
    float lev_st[4];
};

// the decoded prediction value of channel c in [0, 27) of anchor a, from the raw planes
__device__ __forceinline__ float raw_pred(const PostParams& p, int b, int a, int c) {
    int l = 0;
    for (int q = 1; q < p.nlev; ++q) l += a >= p.lev_off[q] ? 1 : 0;
    const int i = a - p.lev_off[l];
    const long long plane = p.lev_off[l + 1] - p.lev_off[l];
    if (c == 26) return p24_sigmoid(p.raw[1][l][b * p.raw_bs[1][l] + i]);
    const float v = p.raw[0][l][b * p.raw_bs[0][l] + c * plane + i];
    if (c >= 2) return expf(v) * p.lev_st[l];                                     // outputs[..., 2:26] = exp(.) * strides
    return (v + (float)(c == 0 ? i % p.lev_w[l] : i / p.lev_w[l])) * p.lev_st[l];  // outputs[..., :2] = (. + grids) * strides
}

struct PostWorkspace {
    size_t tcount, c_score, c_conf, c_anchor, c_cls, c_rect, s_rect, g_keys, g_meta, s_order, s_cell, c_srect, total;
};

inline int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

inline PostWorkspace post_layout(int B, int A) {
    PostWorkspace w;
    size_t off = 0;
    const size_t tiles = (size_t)p24_tiles(A);
    const size_t slots = (size_t)B * tiles * POST_THREADS;
    w.tcount = off;   off = p24_align(off + (size_t)B * tiles * sizeof(int));
    w.c_score = off;  off = p24_align(off + slots * sizeof(float));
    w.c_conf = off;   off = p24_align(off + slots * sizeof(float));
    w.c_anchor = off; off = p24_align(off + slots * sizeof(int));
    w.c_cls = off;    off = p24_align(off + slots * sizeof(int));
    w.c_rect = off;   off = p24_align(off + slots * sizeof(float4));
    w.s_rect = off;   off = p24_align(off + (size_t)B * A * sizeof(float4));
    w.g_keys = off;   off = p24_align(off + (size_t)B * next_pow2(A) * sizeof(unsigned long long));
    w.g_meta = off;   off = p24_align(off + (size_t)B * META_WORDS * sizeof(int));
    w.s_order = off;  off = p24_align(off + (size_t)B * A * sizeof(int));
    w.s_cell = off;   off = p24_align(off + (size_t)B * A * sizeof(int));
    w.c_srect = off;  off = p24_align(off + (size_t)B * A * sizeof(float4));
    w.total = off;
    return w;
}

// ---- mbarrier / TMA bulk copy (1D) ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(phase)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// -------------------------------------------------------------------------------------------
// k_post_filter
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(POST_THREADS) k_post_filter(PostParams p) {
    extern __shared__ __align__(128) float s_rows[];  // [8 warps][32 rows * C]
    __shared__ __align__(8) unsigned long long s_bar[POST_WARPS];
    __shared__ int s_wcnt[POST_WARPS];
    const int b = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int a0 = tile * POST_THREADS + warp * 32;
    const int nrow = max(0, min(32, p.A - a0));
    const int a = a0 + lane;
    const bool active = lane < nrow;
    float* my = s_rows + (size_t)warp * 32 * p.C;
    const float* src = p.pred + (long long)b * p.img_stride + (long long)a0 * p.row_stride;

    // rows of a warp are one contiguous block when row_stride == C; TMA bulk copy needs 16-byte alignment / size
    const unsigned bytes = (unsigned)(nrow * p.C * sizeof(float));
    const bool bulk = nrow > 0 && p.row_stride == p.C && ((uintptr_t)src & 15) == 0 && (bytes & 15) == 0;
    if (bulk) {
        if (lane == 0) {
            mbar_init(&s_bar[warp], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) {
            mbar_expect_tx(&s_bar[warp], bytes);
            bulk_g2s(my, src, bytes, &s_bar[warp]);
        }
        mbar_wait(&s_bar[warp], 0);
    } else {
        for (int r = 0; r < nrow; ++r)
            for (int c = lane; c < p.C; c += 32) my[r * p.C + c] = src[(long long)r * p.row_stride + c];
        __syncwarp();
    }

    // ---- every lane reduces its own row (stride C = 27 + nc floats: odd for nc = 80 -> conflict free) ----------
    bool cand = false;
    float score = 0.f, conf = 0.f;
    int cls = 0;
    float4 rect = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
        const float* row = my + lane * p.C;
        conf = row[27];
        for (int j = 1; j < p.nc; ++j) {  // torch.max: first maximum on ties
            const float v = row[27 + j];
            if (v > conf || (v != v && conf == conf)) {
                conf = v;
                cls = j;
            }
        }
        score = row[26] * conf;          // boxes.py:55, 78
        cand = score >= p.conf_thre;
        if (cand) {
            const float cx = row[0], cy = row[1];
            float x0 = INFINITY, y0 = INFINITY, x1 = -INFINITY, y1 = -INFINITY;
#pragma unroll
            for (int k = 0; k < P24_RAYS; ++k) {
                const float r = row[2 + k];
                const float px = (r * p.coef_x[k]) + cx;  // boxes.py:67-68 (sic: theta * cos theta)
                const float py = (r * p.coef_y[k]) + cy;
                x0 = fminf(x0, px);
                x1 = fmaxf(x1, px);
                y0 = fminf(y0, py);
                y1 = fmaxf(y1, py);
            }
            rect = make_float4(x0, y0, x1, y1);
        }
    }
    // ---- ordered compaction of the tile's candidates -----------------------------------------------------------
    const unsigned bal = __ballot_sync(0xffffffffu, cand);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < POST_WARPS; ++w) {
        const int c = s_wcnt[w];
        base += (w < warp) ? c : 0;
        total += c;
    }
    const long long blk = (long long)b * p.tiles + tile;
    if (cand) {
        const long long o = blk * POST_THREADS + base + __popc(bal & ((1u << lane) - 1u));
        p.c_score[o] = score;
        p.c_conf[o] = conf;
        p.c_anchor[o] = a;
        p.c_cls[o] = cls;
        p.c_rect[o] = rect;
    }
    if (tid == 0) p.tcount[blk] = total;
}

// -------------------------------------------------------------------------------------------
// k_post_filter_raw: the same pass on the head's RAW conv outputs.  One thread per anchor: 32 consecutive anchors of a
// level are 32 consecutive floats of every channel plane, so the planar loads are coalesced.  obj / cls sigmoids and the
// centre / radius decode happen on load, bit for bit like torch (1 / (1 + exp(-x)), (v + grid) * stride, exp(v) * stride).
// class_conf = max_j sigmoid(cls_j), first maximum on ties: the sigmoid is only evaluated for a class whose LOGIT comes
// within a margin of the running best (a smaller logit cannot have a larger sigmoid; equal sigmoids of different logits
// -- the flat ends -- keep the first index because the comparison itself is made on the sigmoids).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(POST_THREADS) k_post_filter_raw(PostParams p) {
    __shared__ int s_wcnt[POST_WARPS];
    const int b = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int a = tile * POST_THREADS + tid;
    const bool active = a < p.A;
    bool cand = false;
    float score = 0.f, conf = 0.f;
    int cls = 0;
    float4 rect = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
        int l = 0;
        for (int q = 1; q < p.nlev; ++q) l += a >= p.lev_off[q] ? 1 : 0;
        const int i = a - p.lev_off[l];
        const long long plane = p.lev_off[l + 1] - p.lev_off[l];
        const float* pc = p.raw[2][l] + b * p.raw_bs[2][l] + i;
        const float obj = p24_sigmoid(p.raw[1][l][b * p.raw_bs[1][l] + i]);
        float xbest = pc[0];
        conf = p24_sigmoid(xbest);
        for (int j0 = 1; j0 < p.nc; j0 += 16) {
            float x[16];  // sixteen planes in flight per thread (the prediction is read once: streaming loads)
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = (j0 + q < p.nc) ? __ldcs(pc + (long long)(j0 + q) * plane) : -INFINITY;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (j0 + q >= p.nc) break;
                const float xv = x[q];
                if (!(xv < xbest - (1e-4f * fabsf(xbest) + 1e-6f))) {  // (also taken for NaN)
                    const float v = p24_sigmoid(xv);
                    if (v > conf || (v != v && conf == conf)) {  // torch.max: first maximum on ties, NaN wins
                        conf = v;
                        cls = j0 + q;
                        xbest = xv;
                    }
                }
            }
        }
        score = obj * conf;              // boxes.py:55, 78
        cand = score >= p.conf_thre;
        if (cand) {
            const float* pr = p.raw[0][l] + b * p.raw_bs[0][l] + i;
            const float st = p.lev_st[l];
            float v[26];
#pragma unroll
            for (int c = 0; c < 26; ++c) v[c] = __ldcs(pr + (long long)c * plane);
            const float cx = (v[0] + (float)(i % p.lev_w[l])) * st, cy = (v[1] + (float)(i / p.lev_w[l])) * st;
            float x0 = INFINITY, y0 = INFINITY, x1 = -INFINITY, y1 = -INFINITY;
#pragma unroll
            for (int k = 0; k < P24_RAYS; ++k) {
                const float r = expf(v[2 + k]) * st;
                const float px = (r * p.coef_x[k]) + cx;  // boxes.py:67-68 (sic: theta * cos theta)
                const float py = (r * p.coef_y[k]) + cy;
                x0 = fminf(x0, px);
                x1 = fmaxf(x1, px);
                y0 = fminf(y0, py);
                y1 = fmaxf(y1, py);
            }
            rect = make_float4(x0, y0, x1, y1);
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, cand);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < POST_WARPS; ++w) {
        const int c = s_wcnt[w];
        base += (w < warp) ? c : 0;
        total += c;
    }
    const long long blk = (long long)b * p.tiles + tile;
    if (cand) {
        const long long o = blk * POST_THREADS + base + __popc(bal & ((1u << lane) - 1u));
        p.c_score[o] = score;
        p.c_conf[o] = conf;
        p.c_anchor[o] = a;
        p.c_cls[o] = cls;
        p.c_rect[o] = rect;
    }
    if (tid == 0) p.tcount[blk] = total;
}

// -------------------------------------------------------------------------------------------
// k_post_nms
// -------------------------------------------------------------------------------------------
// torchvision's devIoU > threshold (boxes without +1, fp32)
__device__ __forceinline__ bool iou_over(const float4 a, const float4 b, float thr) {
    const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
    const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    const float w = fmaxf(right - left, 0.0f), h = fmaxf(bottom - top, 0.0f);
    const float inter = w * h;
    const float sa = (a.z - a.x) * (a.w - a.y);
    const float sb = (b.z - b.x) * (b.w - b.y);
    return (inter / ((sa + sb) - inter)) > thr;
}

// in-place ascending bitonic sort of npad (power of two) 64-bit keys by the whole CTA
__device__ void bitonic_sort(unsigned long long* keys, int npad) {
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {  // one compare-exchange per thread and round
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), ixj = i | j;
                const unsigned long long x = keys[i], y = keys[ixj];
                const bool up = (i & k) == 0;
                if ((x > y) == up) {
                    keys[i] = y;
                    keys[ixj] = x;
                }
            }
            __syncthreads();
        }
    }
}

// One group of batched NMS, by one warp: the group's boxes in sorted order (their sorted ranks in order[g0 .. g0 + m)), the
// classic serial sweep "a kept box suppresses the later boxes it overlaps".  Suppressed boxes get bit 31 of their key.
__device__ __forceinline__ void sweep_group(unsigned long long* keys, const float4* __restrict__ srect,
                                            const int* __restrict__ order, int g0, int m, float thr) {
    const int lane = threadIdx.x & 31;
                    // blocks of 256 boxes, eight per lane in registers: first the kept boxes of the earlier blocks sweep the
                    // block, then the serial sweep inside it
                    for (int blk = 0; blk < m; blk += 256) {
                        const int mb = min(256, m - blk);
                        float4 rc[8];
                        int idx[8];
                        unsigned dead = 0u;
    #pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const int j = t * 32 + lane;
                            idx[t] = j < mb ? order[g0 + blk + j] : -1;
                            rc[t] = j < mb ? srect[idx[t]] : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        for (int i = 0; i < blk; ++i) {
                            const int ii = order[g0 + i];
                            if ((((volatile unsigned long long*)keys)[ii] >> 31) & 1ull) continue;
                            const float4 bi = srect[ii];
    #pragma unroll
                            for (int t = 0; t < 8; ++t) {
                                if (t * 32 < mb && idx[t] >= 0 && !((dead >> t) & 1u)) {
                                    const float4 bj = rc[t];
                                    const bool disjoint = bj.x >= bi.z || bi.x >= bj.z || bj.y >= bi.w || bi.y >= bj.w;
                                    if (!disjoint && iou_over(bi, bj, thr)) dead |= 1u << t;
                                }
                            }
                        }
    #pragma unroll
                        for (int s8 = 0; s8 < 8; ++s8) {
                            if (s8 * 32 < mb) {
                                const int lim = min(32, mb - s8 * 32);
                                for (int l = 0; l < lim; ++l) {
                                    if (__shfl_sync(0xffffffffu, (dead >> s8) & 1u, l)) continue;  // suppressed by a kept box
                                    float4 bi;
                                    bi.x = __shfl_sync(0xffffffffu, rc[s8].x, l);
                                    bi.y = __shfl_sync(0xffffffffu, rc[s8].y, l);
                                    bi.z = __shfl_sync(0xffffffffu, rc[s8].z, l);
                                    bi.w = __shfl_sync(0xffffffffu, rc[s8].w, l);
                                    const int i = s8 * 32 + l;
    #pragma unroll
                                    for (int t = s8; t < 8; ++t) {
                                        const int j = t * 32 + lane;
                                        if (t * 32 < mb && j > i && j < mb && !((dead >> t) & 1u)) {
                                            const float4 bj = rc[t];
                                            const bool disjoint = bj.x >= bi.z || bi.x >= bj.z || bj.y >= bi.w || bi.y >= bj.w;  // IoU = 0
                                            if (!disjoint && iou_over(bi, bj, thr)) dead |= 1u << t;
                                        }
                                    }
                                }
                            }
                        }
    #pragma unroll
                        for (int t = 0; t < 8; ++t)
                            if (idx[t] >= 0 && ((dead >> t) & 1u)) keys[idx[t]] |= (1ull << 31);
                        __syncwarp();
                    }
}

__global__ void __launch_bounds__(NMS_THREADS) k_post_nms(PostParams p) {
    extern __shared__ __align__(16) unsigned long long s_keys[];  // [npad] when the image fits, else unused
    __shared__ int s_prefix[1025];                               // candidates before tile t (tiles <= 1024)
    __shared__ unsigned long long s_mask[64];
    __shared__ float4 s_kept[64];
    __shared__ int s_nkept;
    __shared__ float s_red[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef P24_TIMING
    unsigned long long tm[8];
    int path = 0, dbg_ng = 0, dbg_maxm = 0;
#define PT(k) do { __syncthreads(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm[k])); } while (0)
#else
#define PT(k)
#endif
    PT(0);

    // ---- number of candidates, prefix of the tile counts ----------------------------------------------------------
    if (warp == 0) {  // exclusive scan of the tile counts (tiles <= 1024)
        int carry = 0;
        for (int t0 = 0; t0 < p.tiles; t0 += 32) {
            const int t = t0 + lane;
            const int c = t < p.tiles ? p.tcount[(long long)b * p.tiles + t] : 0;
            int v = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, v, off);
                if (lane >= off) v += u;
            }
            if (t < p.tiles) s_prefix[t] = carry + v - c;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) s_prefix[p.tiles] = carry;
    }
    __syncthreads();
    const int n = s_prefix[p.tiles];
    if (tid == 0) p.cand_count[b] = n;
    if (n == 0) {
        if (tid == 0) p.det_count[b] = 0;
        return;
    }
    int npad = 1;
    while (npad < n) npad <<= 1;
    const bool in_smem = npad <= SORT_SMEM_MAX;
    unsigned long long* keys = in_smem ? s_keys : (p.g_keys + (long long)b * p.npad_global);

    // ---- keys: (descending score, ascending candidate position) -> one ascending 64-bit key; max of the coordinates
    float cmax = -INFINITY;
    const long long slot0 = (long long)b * p.tiles * POST_THREADS;
    for (int t = warp; t < p.tiles; t += NMS_THREADS / 32) {
        const int c = s_prefix[t + 1] - s_prefix[t];
        for (int i = lane; i < c; i += 32) {
            const long long o = slot0 + (long long)t * POST_THREADS + i;
            const unsigned pos = (unsigned)(s_prefix[t] + i);
            keys[pos] = ((unsigned long long)(~p24_ordered(p.c_score[o])) << 32) | pos;
            const float4 r = p.c_rect[o];
            cmax = fmaxf(cmax, fmaxf(fmaxf(r.x, r.y), fmaxf(r.z, r.w)));
            if (p.rect_debug) reinterpret_cast<float4*>(p.rect_debug)[(long long)b * p.A + pos] = r;
        }
    }
    for (int i = n + tid; i < npad; i += NMS_THREADS) keys[i] = ~0ull;
    cmax = warp_max(cmax);
    if (lane == 0) s_red[warp] = cmax;
    __syncthreads();
    cmax = s_red[0];
    for (int w = 1; w < NMS_THREADS / 32; ++w) cmax = fmaxf(cmax, s_red[w]);
    PT(1);
    bitonic_sort(keys, npad);
    PT(2);

    // ---- sorted rectangles (+ class offset for batched NMS: boxes + idxs * (max_coordinate + 1)) --------------------
    float4* srect = p.s_rect + (long long)b * p.A;
    const float step = cmax + 1.0f;
    for (int i = tid; i < n; i += NMS_THREADS) {
        const unsigned pos = (unsigned)(keys[i] & 0xFFFFFFFFull);
        // position -> (tile, rank): binary search in the prefix
        int lo = 0, hi = p.tiles;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_prefix[mid] <= (int)pos) lo = mid;
            else hi = mid;
        }
        const long long o = slot0 + (long long)lo * POST_THREADS + ((int)pos - s_prefix[lo]);
        float4 r = p.c_rect[o];
        if (!p.class_agnostic) {
            const float off = (float)p.c_cls[o] * step;
            r.x = r.x + off;
            r.y = r.y + off;
            r.z = r.z + off;
            r.w = r.w + off;
        }
        srect[i] = r;
        keys[i] = (keys[i] & 0xFFFFFFFF00000000ull) | (unsigned long long)(unsigned)(o - slot0);  // now: slot of the candidate
    }
    __syncthreads();

    // ---- greedy NMS over the sorted boxes.  Bit 31 of a key: suppressed; bit 30: kept (slots < 2^30).
    // Two exact evaluations of the same recursion "a box is kept iff no kept box before it overlaps it":
    //  (I) boxes are binned by x0 into cells at least as wide as the widest box, so that overlapping boxes sit in the
    //      same or in adjacent cells (with the class offsets of batched NMS the cells separate the classes); then
    //      every undecided box looks at the earlier boxes of its three cells: overlapped by a kept one -> suppressed;
    //      every earlier overlapping box suppressed -> kept; else it waits for the next sweep.  The sweeps reach the
    //      fixed point of the recursion in as many rounds as the longest chain of overlaps.
    //  (C) when the boxes do not spread over many cells (class-agnostic NMS): chunks of the next 64 boxes that are
    //      still alive: 64 x 64 tests, a serial pass over the chunk, then the kept ones sweep the boxes after it.
    __shared__ int s_cellstart[NCELL_MAX + 2];
    __shared__ int s_cellfill[NCELL_MAX + 1];
    __shared__ float s_redf[3][32];
    __shared__ int s_cn, s_last;
    __shared__ int s_idx[64];
    float xlo = INFINITY, xhi = -INFINITY, wmx = 0.0f;
    bool finite = true;
    for (int i = tid; i < n; i += NMS_THREADS) {
        const float4 r = srect[i];
        xlo = fminf(xlo, r.x);
        xhi = fmaxf(xhi, r.x);
        wmx = fmaxf(wmx, r.z - r.x);
        finite = finite && (fabsf(r.x) < 1e30f) && (fabsf(r.z) < 1e30f);  // false for NaN / inf
    }
    xlo = -warp_max(-xlo);
    xhi = warp_max(xhi);
    wmx = warp_max(wmx);
    if (lane == 0) {
        s_redf[0][warp] = xlo;
        s_redf[1][warp] = xhi;
        s_redf[2][warp] = wmx;
    }
    const int allfinite = __syncthreads_and(finite ? 1 : 0);
    xlo = s_redf[0][0];
    xhi = s_redf[1][0];
    wmx = s_redf[2][0];
    for (int w = 1; w < NMS_THREADS / 32; ++w) {
        xlo = fminf(xlo, s_redf[0][w]);
        xhi = fmaxf(xhi, s_redf[1][w]);
        wmx = fmaxf(wmx, s_redf[2][w]);
    }
    int ncell = 1;
    float cellw = 1.0f;
    if (allfinite && wmx >= 0.0f) {
        const float span = xhi - xlo;
        cellw = fmaxf(fmaxf(wmx, span * (1.0f / (float)NCELL_MAX)) * 1.0001f, 1e-6f);
        ncell = min(NCELL_MAX, (int)(span / cellw) + 1);
    }
    // (G) batched NMS with a non-negative threshold: boxes of different classes are disjoint after the class offsets (IoU 0),
    //     so the recursion splits into independent GROUPS of classes.  A group is a maximal run of classes whose offset
    //     x ranges [min x0, max x1] overlap (conservative: a class whose boxes reach below x = -1 may touch the class
    //     before it, torchvision lets such boxes interact); almost always one class each.  One warp per group: its boxes
    //     in sorted order, eight per lane in registers, the classic serial sweep "kept box i suppresses the later boxes it
    //     overlaps" with the box broadcast by shuffles.
    PT(3);
    bool done = false, handed = false;
    if (!p.class_agnostic && allfinite && p.nc <= NCELL_MAX / 2 && p.nms_thre >= 0.0f) {
        __shared__ short s_gid[NCELL_MAX / 2];
        __shared__ int s_reach[NCELL_MAX / 2];
        __shared__ unsigned s_pm[2][NCELL_MAX / 2];
        __shared__ int s_ng, s_maxm;
        int* gcnt = s_cellstart;                                   // per group: count, then start (prefix)
        unsigned* hix = reinterpret_cast<unsigned*>(s_cellfill);   // per class: ordered max x1 | ordered max y1 (offset coordinates)
        unsigned* hiy = hix + NCELL_MAX / 2;
        int* grp = p.s_cell + (long long)b * p.A;                  // class, then group, of every sorted box
        int* order = p.s_order + (long long)b * p.A;
        const int nc = p.nc;
        for (int c = tid; c < nc; c += NMS_THREADS) {
            hix[c] = 0u;
            hiy[c] = 0u;
            s_reach[c] = c;
        }
        for (int c = tid; c <= nc; c += NMS_THREADS) gcnt[c] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += NMS_THREADS) {
            const int c = min(max(p.c_cls[slot0 + (long long)(keys[i] & 0x3FFFFFFFull)], 0), nc - 1);
            const float4 r = srect[i];
            atomicMax(&hix[c], p24_ordered(r.z));
            atomicMax(&hiy[c], p24_ordered(r.w));
            grp[i] = c;
        }
        __syncthreads();
        if (warp < 2) {  // prefix maxima over the classes before c (warp 0: x, warp 1: y)
            const unsigned* src = warp == 0 ? hix : hiy;
            unsigned carry = 0u;
            for (int c0 = 0; c0 < nc; c0 += 32) {
                const int c = c0 + lane;
                unsigned v = c < nc ? src[c] : 0u;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, v, off);
                    if (lane >= off) v = max(v, t);
                }
                const unsigned incl = max(v, carry);
                unsigned excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = carry;
                if (c < nc) s_pm[warp][c] = excl;
                carry = __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
        // a box whose upper-left corner lies below the lower-right extent of an earlier class (in x AND in y) may overlap
        // boxes of that class (rare: boxes sticking out of the image corner).  Those few boxes are tested against every box
        // of the earlier classes: only a pair that really exceeds the threshold ties the two classes together.
        __shared__ int s_flag[64];
        __shared__ int s_nflag;
        if (tid == 0) s_nflag = 0;
        __syncthreads();
        for (int i = tid; i < n; i += NMS_THREADS) {
            const int c = grp[i];
            const float4 r = srect[i];
            const unsigned ox = p24_ordered(r.x), oy = p24_ordered(r.y);
            if (ox < s_pm[0][c] && oy < s_pm[1][c]) {
                const int at = atomicAdd(&s_nflag, 1);
                if (at < 64) {
                    s_flag[at] = i;
                } else {  // (too many to test one by one: tie the class to the earliest class it reaches into)
                    for (int c2 = 0; c2 < c; ++c2)
                        if (hix[c2] > ox && hiy[c2] > oy) {
                            atomicMin(&s_reach[c], c2);
                            break;
                        }
                }
            }
        }
        __syncthreads();
        {
            // A partner of a flagged box f overlaps it: its right / lower edge lies beyond f's left / upper edge, which lies at
            // most R = max_f (extent of the earlier classes - f's edge) below that extent: partners are the few boxes whose
            // lower-right corner comes within (Rx, Ry) of their own class's extent.  Flagged x partners, pair by pair.
            __shared__ int s_part[256];
            __shared__ int s_npart;
            __shared__ float s_R[2];
            const int nflag = min(s_nflag, 64);
            if (tid == 0) {
                float rx = 0.0f, ry = 0.0f;
                for (int f = 0; f < nflag; ++f) {
                    const int fi = s_flag[f];
                    const float4 bf = srect[fi];
                    rx = fmaxf(rx, p24_unordered(s_pm[0][grp[fi]]) - bf.x);
                    ry = fmaxf(ry, p24_unordered(s_pm[1][grp[fi]]) - bf.y);
                }
                s_R[0] = rx * 1.001f + 0.05f;  // (slack: the offset coordinates are ~1e5 with an ulp of ~0.01)
                s_R[1] = ry * 1.001f + 0.05f;
                s_npart = 0;
            }
            __syncthreads();
            if (nflag > 0) {
                const float rx = s_R[0], ry = s_R[1];
                for (int i = tid; i < n; i += NMS_THREADS) {
                    const int c = grp[i];
                    const float4 r = srect[i];
                    if (r.z > p24_unordered(hix[c]) - rx && r.w > p24_unordered(hiy[c]) - ry) {
                        const int at = atomicAdd(&s_npart, 1);
                        if (at < 256) s_part[at] = i;
                    }
                }
            }
            __syncthreads();
            const int npart = s_npart;
            if (npart <= 256) {
                for (int q = tid; q < nflag * npart; q += NMS_THREADS) {
                    const int fi = s_flag[q / npart], i = s_part[q % npart];
                    const int cf = grp[fi], c = grp[i];
                    if (c < cf && iou_over(srect[i], srect[fi], p.nms_thre)) atomicMin(&s_reach[cf], c);
                }
            } else {  // (many corner boxes: every flagged box against every box)
                for (int f = 0; f < nflag; ++f) {
                    const int fi = s_flag[f];
                    const int cf = grp[fi];
                    const float4 bf = srect[fi];
                    for (int i = tid; i < n; i += NMS_THREADS) {
                        const int c = grp[i];
                        if (c >= cf) continue;
                        const float4 r = srect[i];
                        if (r.z > bf.x && r.w > bf.y && iou_over(r, bf, p.nms_thre)) atomicMin(&s_reach[cf], c);  // (else disjoint)
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            int sm = nc;
            for (int c = nc - 1; c >= 0; --c) {  // a group starts at c iff no class >= c reaches below c
                sm = min(sm, (int)s_reach[c]);
                s_gid[c] = (short)(sm >= c ? 1 : 0);
            }
            int g = -1;
            for (int c = 0; c < nc; ++c) {
                g += s_gid[c];
                s_gid[c] = (short)g;
            }
            s_ng = g + 1;
        }
        __syncthreads();
        const int ng = s_ng;
        for (int i = tid; i < n; i += NMS_THREADS) {
            const int g = s_gid[grp[i]];
            grp[i] = g;
            atomicAdd(&gcnt[g], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0, mx = 0;
            for (int g = 0; g < ng; ++g) {
                const int c = gcnt[g];
                mx = max(mx, c);
                gcnt[g] = acc;
                acc += c;
            }
            gcnt[ng] = acc;
            s_maxm = mx;
        }
        __syncthreads();
        PT(4);
#ifdef P24_TIMING
        dbg_ng = s_ng;
        dbg_maxm = s_maxm;
#endif
        // largest group that the group sweep handles: beyond it the serial part of a single warp would dominate
        if (s_maxm <= 4096) {
            done = true;
            const float thr = p.nms_thre;
            // ---- every group's boxes in sorted order: a stable counting sort of the sorted ranks by group.  Warp w owns
            // the ranks [w * per, (w + 1) * per): counts per (warp, group), prefix over the warps, placement.
            unsigned short (*s_hist)[128] = reinterpret_cast<unsigned short (*)[128]>(&s_pm[0][0]);  // (s_pm is dead by now)
            static_assert(sizeof(s_pm) >= (NMS_THREADS / 32) * 128 * sizeof(unsigned short), "s_hist does not fit s_pm");
            const bool counting = ng <= 128 && n <= 65535;
            if (counting) {
                const int per = (((n + NMS_THREADS / 32 - 1) / (NMS_THREADS / 32)) + 31) & ~31;
                const int r0 = warp * per, r1 = min(n, r0 + per);
                for (int g = lane; g < ng; g += 32) s_hist[warp][g] = 0;
                __syncwarp();
                for (int i0 = r0; i0 < r1; i0 += 32) {
                    const int i = i0 + lane;
                    const int g = i < r1 ? grp[i] : -1;
                    const unsigned same = __match_any_sync(0xffffffffu, g);
                    if (g >= 0 && lane == __ffs(same) - 1) s_hist[warp][g] += __popc(same);
                    __syncwarp();
                }
                __syncthreads();
                for (int g = tid; g < ng; g += NMS_THREADS) {  // counts -> start of the warp's run inside the group
                    int acc = 0;  // (relative to the group's start: fits 16 bits)
                    for (int w = 0; w < NMS_THREADS / 32; ++w) {
                        const int c = s_hist[w][g];
                        s_hist[w][g] = (unsigned short)acc;
                        acc += c;
                    }
                }
                __syncthreads();
                for (int i0 = r0; i0 < r1; i0 += 32) {
                    const int i = i0 + lane;
                    const int g = i < r1 ? grp[i] : -1;
                    const unsigned same = __match_any_sync(0xffffffffu, g);
                    if (g >= 0) order[gcnt[g] + s_hist[warp][g] + __popc(same & ((1u << lane) - 1u))] = i;
                    __syncwarp();
                    if (g >= 0 && lane == __ffs(same) - 1) s_hist[warp][g] += __popc(same);
                    __syncwarp();
                }
                __syncthreads();
            }
            if (counting) {
                // ---- the sweeps run as a grid of their own over all SMs (k_post_sweep, one warp per group) ----------------
                int* meta = p.g_meta + (long long)b * META_WORDS;
                for (int g = tid; g <= ng; g += NMS_THREADS) meta[2 + g] = gcnt[g];
                if (tid == 0) {
                    meta[0] = 1;
                    meta[1] = ng;
                }
                handed = true;
            }
            // (no lists: too many groups) the warps draw the groups from a counter, the largest first (a group's sweep is
            // serial and grows with the square of its size: the big ones must not come last)
            __shared__ int s_next;
            __syncthreads();
            if (ng <= NCELL_MAX / 2) {
                for (int g = tid; g < ng; g += NMS_THREADS) {
                    const int mg = gcnt[g + 1] - gcnt[g];
                    int rank = 0;
                    for (int g2 = 0; g2 < ng; ++g2) {
                        const int m2 = gcnt[g2 + 1] - gcnt[g2];
                        rank += (m2 > mg || (m2 == mg && g2 < g)) ? 1 : 0;
                    }
                    s_reach[rank] = g;
                }
            }
            if (tid == 0) s_next = handed ? ng : 0;
            __syncthreads();
            for (;;) {
                int gi = 0;
                if (lane == 0) gi = atomicAdd(&s_next, 1);
                gi = __shfl_sync(0xffffffffu, gi, 0);
                if (gi >= ng) break;
                const int g = s_reach[gi];
                const int g0 = gcnt[g], m = gcnt[g + 1] - g0;
                if (m == 0) continue;
                if (!counting) {  // (many groups: one scan of the ranks per group)
                    int pos = g0;
                    for (int i0 = 0; i0 < n && pos < g0 + m; i0 += 32) {
                        const int i = i0 + lane;
                        const bool mine = i < n && grp[i] == g;
                        const unsigned bal = __ballot_sync(0xffffffffu, mine);
                        if (mine) order[pos + __popc(bal & ((1u << lane) - 1u))] = i;
                        pos += __popc(bal);
                    }
                    __syncwarp();
                }
                sweep_group(keys, srect, order, g0, m, thr);
            }
        }
        __syncthreads();
    }
    if (done) {
        // (decided above)
    } else if (ncell >= 16 && p.nms_thre >= 0.0f) {
        // ---- (I) ----
        int* order = p.s_order + (long long)b * p.A;
        int* cellof = p.s_cell + (long long)b * p.A;
        for (int c = tid; c <= ncell; c += NMS_THREADS) {
            s_cellstart[c] = 0;
            s_cellfill[c] = 0;
        }
        __syncthreads();
        for (int i = tid; i < n; i += NMS_THREADS) {
            const int c = min(ncell - 1, max(0, (int)((srect[i].x - xlo) / cellw)));
            cellof[i] = c;
            atomicAdd(&s_cellstart[c + 1], 1);
        }
        __syncthreads();
        if (warp == 0) {  // inclusive scan of the counts -> start of every cell
            int carry = 0;
            for (int c0 = 0; c0 <= ncell; c0 += 32) {
                const int c = c0 + lane;
                int v = c <= ncell ? s_cellstart[c] : 0;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, v, off);
                    if (lane >= off) v += t;
                }
                if (c <= ncell) s_cellstart[c] = v + carry;
                carry += __shfl_sync(0xffffffffu, v, 31);
            }
        }
        __syncthreads();
        float4* crect = p.c_srect + (long long)b * p.A;
        for (int i = tid; i < n; i += NMS_THREADS) {
            const int c = cellof[i];
            const int e = s_cellstart[c] + atomicAdd(&s_cellfill[c], 1);
            order[e] = i;
            crect[e] = srect[i];
        }
        __syncthreads();
        // Warp w owns the cells c with c % 32 == w (adjacent cells belong to different warps) and walks through ITS boxes
        // in sorted order, one box at a time, the lanes over the earlier boxes of the box's three cells.  An earlier
        // overlapping box of another warp that is not decided yet is waited for: every wait points to a smaller
        // sorted index and every warp advances in sorted order, so the smallest undecided box can always be decided.
        for (int base = 0; base < n; base += 32) {
            const int im = base + lane;
            const bool mine = im < n && (cellof[im] & 31) == warp;
            unsigned todo = __ballot_sync(0xffffffffu, mine);
            while (todo) {
                const int i = base + __ffs(todo) - 1;
                todo &= todo - 1;
                const float4 bi = srect[i];
                const int ci = cellof[i];
                const int e0 = s_cellstart[max(ci - 1, 0)], e1 = s_cellstart[min(ci + 2, ncell)];
                bool dead = false;
                // the entries of the three cells are contiguous in cell order: coalesced loads, the next 32 entries
                // are requested before the current ones are tested
                int jn = (e0 + lane < e1) ? order[e0 + lane] : 0x7fffffff;
                float4 rn = (e0 + lane < e1) ? crect[e0 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int eb = e0; eb < e1 && !dead; eb += 32) {
                    const int j = jn;
                    const float4 bj = rn;
                    const int en = eb + 32 + lane;
                    jn = (en < e1) ? order[en] : 0x7fffffff;
                    if (en < e1) rn = crect[en];
                    bool hit = false;
                    {
                        if (j < i) {
                            const bool disjoint = bj.x >= bi.z || bi.x >= bj.z || bj.y >= bi.w || bi.y >= bj.w;  // IoU = 0
                            if (!disjoint && iou_over(bj, bi, p.nms_thre)) {
                                unsigned st;
                                do {
                                    st = (unsigned)(((volatile unsigned long long*)keys)[j] >> 30) & 3u;
                                } while (st == 0u);
                                hit = (st & 1u) != 0u;  // overlapped by a kept box
                            }
                        }
                    }
                    dead = __any_sync(0xffffffffu, hit);
                }
                if (lane == 0) {
                    atomicOr(&keys[i], dead ? (1ull << 31) : (1ull << 30));
                    __threadfence_block();
                }
                __syncwarp();
            }
        }
    } else {
        // ---- (C) ----
        int cursor = 0;
        while (cursor < n) {
            // the next (up to) 64 boxes that are still alive, in order
            if (tid == 0) {
                s_cn = 0;
                s_last = n;
            }
            __syncthreads();
            for (int w0 = cursor; w0 < n; w0 += NMS_THREADS) {
                const int i = w0 + tid;
                const bool alive = i < n && !((keys[i] >> 31) & 1ull);
                const unsigned bal = __ballot_sync(0xffffffffu, alive);
                if (lane == 0) s_red[warp] = __int_as_float(__popc(bal));
                __syncthreads();
                int before = s_cn;
                for (int w = 0; w < warp; ++w) before += __float_as_int(s_red[w]);
                int tot = 0;
                for (int w = 0; w < NMS_THREADS / 32; ++w) tot += __float_as_int(s_red[w]);
                if (alive) {
                    const int k = before + __popc(bal & ((1u << lane) - 1u));
                    if (k < 64) s_idx[k] = i;
                    if (k == 63) s_last = i + 1;
                }
                __syncthreads();
                if (tid == 0) s_cn = min(64, s_cn + tot);
                __syncthreads();
                if (s_cn >= 64) break;
            }
            const int cn = s_cn;
            const int cend = s_last;  // everything before cend is decided after this round
            if (cn == 0) break;
            // row i of the 64 x 64 overlap matrix by one warp: two ballots, no atomics (64-bit shared-memory atomics are
            // compare-and-swap loops)
            for (int i = warp; i < cn; i += NMS_THREADS / 32) {
                const float4 bi = srect[s_idx[i]];
                const int j0 = lane, j1 = lane + 32;
                const bool t0 = j0 > i && j0 < cn && iou_over(bi, srect[s_idx[j0]], p.nms_thre);
                const bool t1 = j1 > i && j1 < cn && iou_over(bi, srect[s_idx[j1]], p.nms_thre);
                const unsigned lo = __ballot_sync(0xffffffffu, t0), hi = __ballot_sync(0xffffffffu, t1);
                if (lane == 0) s_mask[i] = (unsigned long long)lo | ((unsigned long long)hi << 32);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned long long removed = 0ull;
                int nk = 0;
                for (int i = 0; i < cn; ++i) {
                    if ((removed >> i) & 1ull) {
                        keys[s_idx[i]] |= (1ull << 31);
                    } else {
                        removed |= s_mask[i];
                        s_kept[nk++] = srect[s_idx[i]];
                    }
                }
                s_nkept = nk;
            }
            __syncthreads();
            const int nk = s_nkept;
            for (int j = cend + tid; j < n; j += NMS_THREADS) {
                if ((keys[j] >> 31) & 1ull) continue;
                const float4 bj = srect[j];
                bool dead = false;
                for (int i = 0; i < nk && !dead; ++i) dead = iou_over(s_kept[i], bj, p.nms_thre);
                if (dead) keys[j] |= (1ull << 31);
            }
            __syncthreads();
            cursor = cend;
        }
    }
    __syncthreads();

    PT(5);
#ifdef P24_TIMING
    path = done ? 1 : 2;
#endif
    // ---- hand over: the keys (slot + kept / suppressed bits) in global memory, the sweeps pending or not -----------------
    if (!handed && tid == 0) p.g_meta[(long long)b * META_WORDS] = 0;
    if (in_smem) {
        unsigned long long* gk = p.g_keys + (long long)b * p.npad_global;
        for (int i = tid; i < n; i += NMS_THREADS) gk[i] = keys[i];
    }
#ifdef P24_TIMING
    PT(6);
    if (tid == 0 && (b == 0 || b == 7))
        printf("ng=%d maxm=%d nms b=%d n=%d path=%d handed=%d keys %.1f sort %.1f rects %.1f groups %.1f nms+lists %.1f store %.1f us\n",
               dbg_ng, dbg_maxm, b, n, path, (int)handed, (tm[1] - tm[0]) * 1e-3, (tm[2] - tm[1]) * 1e-3, (tm[3] - tm[2]) * 1e-3,
               (tm[4] - tm[3]) * 1e-3, (tm[5] - tm[4]) * 1e-3, (tm[6] - tm[5]) * 1e-3);
#endif
}

// k_post_sweep: the group sweeps of batched NMS as a grid over all SMs: one warp per group (ceil(groups / 8) x B CTAs)
__global__ void __launch_bounds__(256) k_post_sweep(PostParams p) {
    const int b = blockIdx.y, warp = threadIdx.x >> 5;
    const int* meta = p.g_meta + (long long)b * META_WORDS;
    if (p.cand_count[b] == 0 || meta[0] != 1) return;
    const int g = blockIdx.x * 8 + warp;
    if (g >= meta[1]) return;
    const int g0 = meta[2 + g], m = meta[3 + g] - g0;
    if (m <= 0) return;
    sweep_group(p.g_keys + (long long)b * p.npad_global, p.s_rect + (long long)b * p.A, p.s_order + (long long)b * p.A, g0, m,
                p.nms_thre);
}

// k_post_out: OUT_SPLIT CTAs per image: the kept boxes of a quarter of the sorted ranks -> output rows in sorted order
__global__ void __launch_bounds__(NMS_THREADS) k_post_out(PostParams p) {
    __shared__ float s_red[32];
    __shared__ int s_cnt[32];
    const int b = blockIdx.y, q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.cand_count[b];
    const unsigned long long* keys = p.g_keys + (long long)b * p.npad_global;
    const long long slot0 = (long long)b * p.tiles * POST_THREADS;
    const int chunk = (((n + OUT_SPLIT - 1) / OUT_SPLIT) + NMS_THREADS - 1) / NMS_THREADS * NMS_THREADS;
    const int r0 = min(n, q * chunk), r1 = min(n, r0 + chunk);
    // kept boxes before my range
    int before0 = 0;
    for (int i = tid; i < r0; i += NMS_THREADS) before0 += ((keys[i] >> 31) & 1ull) ? 0 : 1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) before0 += __shfl_xor_sync(0xffffffffu, before0, off);
    if (lane == 0) s_cnt[warp] = before0;
    __syncthreads();
    before0 = 0;
    for (int w = 0; w < NMS_THREADS / 32; ++w) before0 += s_cnt[w];
    __syncthreads();
    // ---- output rows of the survivors in sorted order ------------------------------------------------------------
    // ordered compaction over i = 0..n-1
    __shared__ int s_base;
    if (tid == 0) s_base = before0;
    __syncthreads();
    int* klist = p.s_order + (long long)b * p.A;  // sorted rank of the k-th kept box (the group lists are no longer needed)
    for (int i0 = r0; i0 < r1; i0 += NMS_THREADS) {
        const int i = i0 + tid;
        const bool keep = i < r1 && !((keys[i] >> 31) & 1ull);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_red[warp] = __int_as_float(__popc(bal));
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += __float_as_int(s_red[w]);
        int tot = 0;
        for (int w = 0; w < NMS_THREADS / 32; ++w) tot += __float_as_int(s_red[w]);
        if (keep) klist[before + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (tid == 0) s_base += tot;
        __syncthreads();
    }
    // rows: one warp per kept box, the 27 leading floats of its prediction row in one coalesced read
    const int nkept = s_base;
    for (int k0 = before0 + warp * 32; k0 < nkept; k0 += NMS_THREADS) {
        // the lanes fetch what 32 rows need (rank -> slot -> anchor, confidence, class), then the warp copies the rows,
        // four in flight
        const int kk = k0 + lane;
        int a = 0;
        float conf = 0.0f, clsf = 0.0f;
        if (kk < nkept) {
            const int i = klist[kk];
            const long long o = slot0 + (long long)(keys[i] & 0x3FFFFFFFull);
            a = p.c_anchor[o];
            conf = p.c_conf[o];
            clsf = (float)p.c_cls[o];
            p.keep_idx[(long long)b * p.A + kk] = a;
        }
        const int cnt = min(32, nkept - k0);
        for (int r0 = 0; r0 < cnt; r0 += 4) {
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = min(r0 + q, cnt - 1);
                const int ar = __shfl_sync(0xffffffffu, a, r);
                const float cr = __shfl_sync(0xffffffffu, conf, r), lr = __shfl_sync(0xffffffffu, clsf, r);
                if (p.pred) {
                    const float* row = p.pred + (long long)b * p.img_stride + (long long)ar * p.row_stride;
                    v[q] = lane < 27 ? row[lane] : (lane == 27 ? cr : lr);
                } else {
                    v[q] = lane < 27 ? raw_pred(p, b, ar, lane) : (lane == 27 ? cr : lr);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (r0 + q < cnt && lane < 29) p.det_rows[((long long)b * p.A + k0 + r0 + q) * 29 + lane] = v[q];
        }
    }
    if (tid == 0 && q == OUT_SPLIT - 1) p.det_count[b] = s_base;
}

}  // namespace

extern "C" size_t p24_postprocess_workspace_bytes(int B, int A) {
    if (B <= 0 || A <= 0) return 0;
    return post_layout(B, A).total;
}

namespace {
int post_impl(const float* prediction, int64_t img_stride, int64_t row_stride, const float* const* h_raw,
              const int64_t* h_raw_bs, const int32_t* h_levels, int n_levels, int B, int A,
              int num_classes, const float* h_coef_x, const float* h_coef_y, float conf_thre,
              float nms_thre, int class_agnostic, int32_t* cand_count, int32_t* det_count,
              float* det_rows, int32_t* keep_idx, float* rect_debug, void* workspace,
              size_t workspace_bytes, void* stream) {
    if ((!prediction && !h_raw) || !h_coef_x || !h_coef_y || !cand_count || !det_count || !det_rows || !keep_idx || !workspace)
        return P24_E_BADARG;
    if (B <= 0 || A <= 0 || num_classes <= 0) return P24_E_BADARG;
    if (((uintptr_t)workspace & 255) != 0) return P24_E_BADARG;
    const PostWorkspace L = post_layout(B, A);
    if (workspace_bytes < L.total) return P24_E_WORKSPACE;
    const int tiles = p24_tiles(A);
    if (tiles > 1024) return P24_E_UNSUPPORTED;
    const int C = 27 + num_classes;
    const size_t smem_filter = (size_t)POST_WARPS * 32 * C * sizeof(float);
    if (prediction && smem_filter > 200 * 1024) return P24_E_UNSUPPORTED;
    char* ws = (char*)workspace;
    PostParams p;
    memset(&p, 0, sizeof(p));
    if (!prediction) {
        if (!h_raw_bs || !h_levels || n_levels < 1 || n_levels > 4) return P24_E_BADARG;
        int total = 0;
        p.nlev = n_levels;
        for (int l = 0; l < n_levels; ++l) {
            const int off = h_levels[4 * l], W = h_levels[4 * l + 1], H = h_levels[4 * l + 2];
            if (off != total || W <= 0 || H <= 0) return P24_E_BADARG;
            p.lev_off[l] = off;
            p.lev_w[l] = W;
            memcpy(&p.lev_st[l], &h_levels[4 * l + 3], sizeof(float));
            total += W * H;
            for (int t = 0; t < 3; ++t) {
                p.raw[t][l] = h_raw[t * n_levels + l];
                p.raw_bs[t][l] = h_raw_bs[t * n_levels + l];
                if (!p.raw[t][l]) return P24_E_BADARG;
            }
        }
        if (total != A) return P24_E_BADARG;
        p.lev_off[n_levels] = A;
    }
    p.pred = prediction; p.img_stride = img_stride; p.row_stride = row_stride;
    p.B = B; p.A = A; p.nc = num_classes; p.C = C;
    for (int k = 0; k < P24_RAYS; ++k) {
        p.coef_x[k] = h_coef_x[k];
        p.coef_y[k] = h_coef_y[k];
    }
    p.conf_thre = conf_thre; p.nms_thre = nms_thre; p.class_agnostic = class_agnostic;
    p.cand_count = cand_count; p.det_count = det_count; p.det_rows = det_rows; p.keep_idx = keep_idx;
    p.rect_debug = rect_debug;
    p.tcount = (int*)(ws + L.tcount);
    p.c_score = (float*)(ws + L.c_score);
    p.c_conf = (float*)(ws + L.c_conf);
    p.c_anchor = (int*)(ws + L.c_anchor);
    p.c_cls = (int*)(ws + L.c_cls);
    p.c_rect = (float4*)(ws + L.c_rect);
    p.s_rect = (float4*)(ws + L.s_rect);
    p.g_keys = (unsigned long long*)(ws + L.g_keys);
    p.g_meta = (int*)(ws + L.g_meta);
    p.s_order = (int*)(ws + L.s_order);
    p.s_cell = (int*)(ws + L.s_cell);
    p.c_srect = (float4*)(ws + L.c_srect);
    p.tiles = tiles;
    p.npad_global = next_pow2(A);
    cudaStream_t st = (cudaStream_t)stream;
    const int npad_max = next_pow2(A);
    const size_t smem_nms = (size_t)(npad_max <= SORT_SMEM_MAX ? npad_max : 0) * sizeof(unsigned long long);
    if (p24::dev_once(1u << 8)) {  // per device: a process may drive several GPUs
        cudaFuncSetAttribute(k_post_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_post_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM_MAX * 8);
    }
    p24::prof_mark(4, st);
    if (prediction) k_post_filter<<<dim3(tiles, B), POST_THREADS, smem_filter, st>>>(p);
    else k_post_filter_raw<<<dim3(tiles, B), POST_THREADS, 0, st>>>(p);
    p24::prof_mark(5, st);
    k_post_nms<<<B, NMS_THREADS, smem_nms, st>>>(p);  // sort, groups, lists (+ the NMS itself on the other paths)
    {
        const int gmax = num_classes < SPLIT_GROUPS ? num_classes : SPLIT_GROUPS;
        if (!class_agnostic) k_post_sweep<<<dim3((gmax + 7) / 8, B), 256, 0, st>>>(p);
        k_post_out<<<dim3(OUT_SPLIT, B), NMS_THREADS, 0, st>>>(p);
    }
    p24::prof_mark(6, st);
    return (int)cudaGetLastError();
}
}  // namespace

extern "C" int p24_postprocess(const float* prediction, int64_t img_stride, int64_t row_stride, int B, int A,
                               int num_classes, const float* h_coef_x, const float* h_coef_y, float conf_thre,
                               float nms_thre, int class_agnostic, int32_t* cand_count, int32_t* det_count,
                               float* det_rows, int32_t* keep_idx, float* rect_debug, void* workspace,
                               size_t workspace_bytes, void* stream) {
    if (!prediction) return P24_E_BADARG;
    return post_impl(prediction, img_stride, row_stride, nullptr, nullptr, nullptr, 0, B, A, num_classes, h_coef_x, h_coef_y,
                     conf_thre, nms_thre, class_agnostic, cand_count, det_count, det_rows, keep_idx, rect_debug, workspace,
                     workspace_bytes, stream);
}

extern "C" int p24_postprocess_raw(const float* const* h_raw, const int64_t* h_raw_batch_stride, const int32_t* h_levels,
                                   int n_levels, int B, int A, int num_classes, const float* h_coef_x, const float* h_coef_y,
                                   float conf_thre, float nms_thre, int class_agnostic, int32_t* cand_count,
                                   int32_t* det_count, float* det_rows, int32_t* keep_idx, float* rect_debug, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    if (!h_raw) return P24_E_BADARG;
    return post_impl(nullptr, 0, 0, h_raw, h_raw_batch_stride, h_levels, n_levels, B, A, num_classes, h_coef_x, h_coef_y,
                     conf_thre, nms_thre, class_agnostic, cand_count, det_count, det_rows, keep_idx, rect_debug, workspace,
                     workspace_bytes, stream);
}
