// p24_simota.cu — the fused YOLOX-24p SimOTA assignment + loss-sum path for sm_100a.
//
// Replaces, for a whole batch and without ever writing the G x A cost matrix to memory:
//   Loss_Function.get_assignments / get_in_boxes_info / pts_in_poly   models/losses.py:359-592
//   utils.boxes.bboxes_iou + pairwise circle_inter                   utils/boxes.py:102-243
//   Loss_Function.dynamic_k_matching                                 models/losses.py:444-494
//   the loss sums and re-weighting of Loss_Function.forward          models/losses.py:246-345
//
// Kernel chain (caller's stream, no host synchronisation, programmatic dependent launch between stages):
//   k_pass         two kinds of CTA side by side, both self-sufficient (each computes nlabel and the per-GT records
//                  of its image from the label rows):
//                  - anchor CTAs, one per 256-anchor tile: the ONE pass over the head output (cp.async reads of the
//                    27 geometry channels of every row).  Candidate mask (polygon test OR centre window) with
//                    geometric pruning and an atan2-free angle test, the compacted candidate list, per-tile seed
//                    values for the top-10 bracket, sum BCEWithLogits(obj, 0), outputs initialised to background;
//                  - centre-window CTAs: the (GT, centre-window anchor) pairs enumerated from the level grids:
//                    polygon test, exact pair value and SimOTA cost -> the GT's window cost table
//   k_match        one CTA per GT: dynamic k from the top-10-largest pair values over the candidates (bracketed by
//                  seed values and a monotone upper bound; bound-filtered exact evaluation otherwise), then the k
//                  smallest costs of the window table -> claims (spill into the penalised regime when there are
//                  too few valid anchors)
//   k_resolve_loss one warp per claimed anchor: conflict resolution (argmin over all GTs), fg_mask / matched_gt /
//                  pred_iou, the 28 loss sums; the last CTA reduces them in a fixed order and, when asked,
//                  applies the normalisation and the stateful re-weighting (losses.py:280-345)
//
// Compile with -fmad=false: the discrete decisions hang on fp32 thresholds evaluated in the
// reference's operation order; bounds and fast paths use explicit fmaf.
#include "p24_common.cuh"
#include "p24_host.h"
#include <string.h>

namespace {

struct Level {
    int off, W, H, pad;  // anchors [off, off + W * H) form a W x H grid, row-major (yolo_head_24p.py:222-230)
};

struct Params {
    const float* outputs;
    long long img_stride, row_stride;
    int B, A, nc;
    const float* labels;
    long long lab_img_stride, lab_row_stride;
    int Lmax;
    const float* x_shifts;
    const float* y_shifts;
    const float* strides;
    uint8_t* fg_mask;
    int32_t* matched_gt;
    float* pred_iou;
    int32_t* num_fg;
    int32_t* num_gt;
    int32_t* dyn_k;
    float* sums28;
    float* state26;  // optional fused finalize
    float* result54;
    float* weights27;
    // workspace
    float* gt_rec;
    float4* clist;
    float* sval;
    float* wtab;
    float* tbox;
    float* seg;       // [B, tiles, 8 warps, 8]: per-warp boxes / counts of the candidate lists
    int* ccount;
    int* claim_cnt;
    int* claim_gt;
    double* obj_part;
    int* claimed;
    int* nclaimed;
    long long* acc_fix;
    unsigned* ticket;
    int* err_flag;
    unsigned flags;
    float* mbox[P24_MAX_RANKS];  // peer mailboxes of the fused all-reduce (nranks > 1)
    int rank, nranks;
    unsigned epoch;
    int tiles;
    int nlev;
    Level lev[P24_MAX_LEVELS];
};

// Debug-only phase timers (-DP24_TIMING): thread 0 of every CTA stores %globaltimer at phase boundaries.
#ifdef P24_TIMING
__device__ unsigned long long g_tstamp[6][4096][20];
__device__ __forceinline__ void tmark(int kern, int cta, int slot) {
    if (threadIdx.x == 0 && cta >= 0 && cta < 4096) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_tstamp[kern][cta][slot] = t;
    }
}
#define TMARK(kern, cta, slot) tmark(kern, cta, slot)
#else
#define TMARK(kern, cta, slot)
#endif

#define MATCH_THREADS 256
#define MATCH_WARPS (MATCH_THREADS / 32)
#define MATCH_GROUPS (MATCH_THREADS / 8)

__device__ __forceinline__ void pdl_wait() {
#if __CUDA_ARCH__ >= 900
    cudaGridDependencySynchronize();
#endif
}
// lets the next kernel of the stream start launching now (it must not read this kernel's outputs before its own
// pdl_wait)
__device__ __forceinline__ void pdl_trigger() {
#if __CUDA_ARCH__ >= 900
    cudaTriggerProgrammaticLaunchCompletion();
#endif
}

// ---- 8-lane group reductions (the group's own mask: groups of a warp may diverge) -----------------------
__device__ __forceinline__ unsigned group_mask() { return 0xFFu << (threadIdx.x & 24); }
__device__ __forceinline__ float group_sum(float v, unsigned m) {
    v += __shfl_xor_sync(m, v, 1);
    v += __shfl_xor_sync(m, v, 2);
    v += __shfl_xor_sync(m, v, 4);
    return v;
}
__device__ __forceinline__ float group_prod(float v, unsigned m) {
    v *= __shfl_xor_sync(m, v, 1);
    v *= __shfl_xor_sync(m, v, 2);
    v *= __shfl_xor_sync(m, v, 4);
    return v;
}
__device__ __forceinline__ int group_sum_i(int v, unsigned m) {
    v += __shfl_xor_sync(m, v, 1);
    v += __shfl_xor_sync(m, v, 2);
    v += __shfl_xor_sync(m, v, 4);
    return v;
}

// -------------------------------------------------------------------------------------------
// GT preparation, executed inside every CTA of k_pass for its own image (no separate kernel: the records are a few
// hundred instructions per GT and the label rows stay in L2)
// -------------------------------------------------------------------------------------------
// nlabel = (labels.sum(2) > 0).sum(1)   losses.py:190 ; the first n rows are the GTs (losses.py:219-220).
// Four threads per label row (double partial sums); every thread returns the count.  Contains __syncthreads().
__device__ __forceinline__ int block_count_labels(const Params& p, const float* __restrict__ lab, int* s_n) {
    const int tid = threadIdx.x;
    if (tid == 0) *s_n = 0;
    __syncthreads();
    if (!(p.flags & P24_F_ALL_ROWS)) {
        const int part = tid & 3;
        int local = 0;
        for (int r0 = 0; r0 < p.Lmax; r0 += P24_THREADS / 4) {
            const int r = r0 + (tid >> 2);
            double s = 0.0;
            if (r < p.Lmax) {
                const float* row = lab + (long long)r * p.lab_row_stride;
                const int c0 = part * 13, c1 = min(51, c0 + 13);
                for (int c = c0; c < c1; ++c) s += (double)row[c];
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            local += (part == 0 && r < p.Lmax && (float)s > 0.0f) ? 1 : 0;
        }
        local = warp_sum_i(local);
        if ((tid & 31) == 0 && local) atomicAdd(s_n, local);
    }
    __syncthreads();
    return (p.flags & P24_F_ALL_ROWS) ? p.Lmax : *s_n;
}

// the record of one GT (p24_common.cuh) from its label row, by one warp (lanes over the 24 vertices)
__device__ __forceinline__ void warp_gt_record(const float* __restrict__ row, float* __restrict__ rec) {
    const int lane = threadIdx.x & 31;
    const float cx = row[1], cy = row[2];
    const int k = lane < P24_RAYS ? lane : 0;
    const int k2 = (k == P24_RAYS - 1) ? 0 : k + 1;
    const float x = row[3 + 2 * k], y = row[4 + 2 * k];
    const float x2 = row[3 + 2 * k2], y2 = row[4 + 2 * k2];
    const float rg = p24_gt_radius(x - cx, y - cy);
    const float ex = x2 - x, ey = y2 - y;
    const float len2 = fmaf(ex, ex, ey * ey);
    float len = sqrtf(len2);
    // distance from the centre to the edge segment
    const float wx = cx - x, wy = cy - y;
    float tt = len2 > 0.0f ? __fdividef(fmaf(wx, ex, wy * ey), len2) : 0.0f;
    tt = fminf(fmaxf(tt, 0.0f), 1.0f);
    const float qx = wx - tt * ex, qy = wy - tt * ey;
    float rin = sqrtf(fmaf(qx, qx, qy * qy));
    // crossing-number parity of the centre
    bool cross = false;
    if ((y > cy) != (y2 > cy)) {
        const float xi = fmaf(ex, __fdividef(cy - y, ey), x);
        cross = cx < xi;
    }
    float rgmax = rg, rgmin = rg;
    if (lane >= P24_RAYS) {
        len = 0.0f;
        rin = INFINITY;
        cross = false;
        rgmax = 0.0f;
        rgmin = INFINITY;
    }
    const unsigned par = __ballot_sync(0xffffffffu, cross);
    const bool nan_any = __any_sync(0xffffffffu, !(rin == rin) && lane < P24_RAYS);
    const float perim = warp_sum(len);
    const float rgsum = warp_sum(lane < P24_RAYS ? rg : 0.0f);
    const float rg2sum = warp_sum(lane < P24_RAYS ? rg * rg : 0.0f);
    rgmax = warp_max(rgmax);
    rgmin = -warp_max(-rgmin);
    rin = -warp_max(-rin);
    if (lane < P24_RAYS) {
        rec[GT_VX + lane] = x;
        rec[GT_VY + lane] = y;
        rec[GT_RG + lane] = rg;
    }
    if (lane == 0) {
        const bool inside = (__popc(par) & 1) != 0;
        // A point inside a closed polygon has |winding| >= 1, so its total unsigned angle is >= 360 degrees:
        // a disc around an interior centre that stays clear of every edge passes the >= 350 test (2 % + 0.01 px
        // of slack covers the fp32 evaluation of the distances above).
        float ra = (inside && !nan_any) ? fmaf(0.98f, rin, -0.01f) : 0.0f;
        ra = fmaxf(ra, 0.0f);
        // Outside, the angle sum is <= perimeter / distance-to-polygon (radians): it is < 349 degrees beyond
        // rgmax + perimeter * (180/pi) / 349 (1 % slack).
        const float rr = fmaf(perim * 1.01f, 57.29578f / 349.0f, rgmax) * 1.001f + 1e-2f;
        float rrej2 = rr * rr;
        if (!(rrej2 == rrej2)) rrej2 = INFINITY;  // NaN labels: never reject
        rec[GT_CX] = cx;
        rec[GT_CY] = cy;
        rec[GT_RIN2] = ra * ra;
        rec[GT_RREJ2] = rrej2;
        rec[GT_CLS] = row[0];
        rec[GT_RGMAX] = rgmax;
        rec[GT_RGMIN] = rgmin;
        rec[7] = 0.0f;
        rec[GT_RGMS] = rg2sum * (1.0f / 24.0f);
        rec[GT_RGMEAN] = rgsum * (1.0f / 24.0f);
        rec[58] = 0.0f;
        rec[59] = 0.0f;
    }
}

// k_gt_prep: one CTA per image.  It lets k_pass launch at once (programmatic dependent launch): k_pass's CTAs stage
// their first rows while this kernel runs and wait for it only before they read the records.
#define PREP_THREADS 256
__global__ void __launch_bounds__(PREP_THREADS) k_gt_prep(const __grid_constant__ Params p) {
    // launched as a programmatic dependent of whatever precedes it in the stream (in back-to-back steps: the previous
    // step's k_resolve_loss, which triggers at once): resident early, it starts the moment that work is complete.  It
    // must not let k_pass go before that: k_pass draws tickets that the previous step's last CTA resets.
    pdl_wait();
    pdl_trigger();
    TMARK(3, blockIdx.x, 0);
    __shared__ int s_n;
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
    const float* lab = p.labels + (long long)b * p.lab_img_stride;
    const int n = block_count_labels(p, lab, &s_n);
    if (tid == 0) {
        p.num_gt[b] = n;
        p.num_fg[b] = 0;
        p.nclaimed[b] = 0;
        atomicMax(&p.ticket[2 + p.B], (unsigned)n);  // the batch's largest num_gt: k_pass lays its window items out for it
    }
    for (int g = warp; g < n; g += PREP_THREADS / 32)
        warp_gt_record(lab + (long long)g * p.lab_row_stride, p.gt_rec + ((long long)b * p.Lmax + g) * GT_REC);
    TMARK(3, blockIdx.x, 1);
}

// -------------------------------------------------------------------------------------------
// shared device helpers
// -------------------------------------------------------------------------------------------
// Out-of-line copies of the two heavy scalar routines: the kernels call them from many places, and inlining every
// call made the per-GT kernel 17k instructions long (instruction-cache bound).
__device__ __noinline__ float ray_loss(float rg, float rp, float d) { return p24_ray_loss(rg, rp, d); }
__device__ __noinline__ float edge_angle(float sx, float sy, float ex, float ey) { return p24_edge_angle(sx, sy, ex, ey); }

// exact pair value of (GT record, prediction row in global memory): utils/boxes.py:166-243, one thread
__device__ __noinline__ float pair_value_row(const float* __restrict__ rec, const float* __restrict__ row) {
    const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], row[0], row[1]);
    float s = 0.0f;
#pragma unroll 1
    for (int k = 0; k < P24_RAYS; ++k) s = s + ray_loss(rec[GT_RG + k], row[2 + k], d);
    return (s / 24.0f) / 2.0f;
}

// the same value by an 8-lane group (3 rays per lane, fixed reduction tree); every lane of the group returns it
__device__ __forceinline__ float group_pair_value(const float* __restrict__ rec, const float* __restrict__ row, unsigned m) {
    const int sub = threadIdx.x & 7;
    const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], row[0], row[1]);
    float s = 0.0f;
#pragma unroll 1
    for (int q = 0; q < 3; ++q) {
        const int k = sub * 3 + q;
        s = s + ray_loss(rec[GT_RG + k], row[2 + k], d);
    }
    s = group_sum(s, m);
    return (s / 24.0f) / 2.0f;
}

// A certified LOWER bound of the pair value by an 8-lane group, for seeds: when every ray is in the "apart" branch
// (d >= rg + rp, the same fp32 comparison as the reference) the value has the closed form
// (1/48) sum_k (2 - 4 (rg^2 + rp^2) / (rg + rp + d)^2), evaluated here in fast arithmetic; it differs from the
// reference-order fp32 value by < 3e-6, so value - 1e-5 is a valid lower bound.  Other pairs are evaluated exactly.
__device__ __forceinline__ float group_pair_value_lb(const float* __restrict__ rec, const float* __restrict__ row, unsigned m) {
    const int sub = threadIdx.x & 7;
    const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], row[0], row[1]);
    float rg[3], rp[3];
    bool apart = true;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        rg[q] = rec[GT_RG + sub * 3 + q];
        rp[q] = row[2 + sub * 3 + q];
        apart = apart && (d >= rg[q] + rp[q]);
    }
    const unsigned all = __ballot_sync(m, apart);
    if ((all & m) == m) {
        float s = 0.0f;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float t = (rg[q] + rp[q]) + d;
            s += 2.0f - __fdividef(4.0f * fmaf(rg[q], rg[q], rp[q] * rp[q]), t * t);
        }
        s = group_sum(s, m);
        return s * (1.0f / 48.0f) - 1e-5f;
    }
    float s = 0.0f;
#pragma unroll 1
    for (int q = 0; q < 3; ++q) s = s + ray_loss(rg[q], rp[q], d);
    s = group_sum(s, m);
    return (s / 24.0f) / 2.0f;
}

__device__ __forceinline__ int gt_class(const float* rec, int nc) {
    const int c = (int)rec[GT_CLS];
    return min(max(c, 0), nc - 1);
}

// Sum over all classes of BCE(p_j, 0) (losses.py:406-416) in product form, by the 32 lanes of a warp
__device__ __noinline__ float warp_cls_neg_sum(const float* __restrict__ cls, int nc, float eo1) {
    const int lane = threadIdx.x & 31;
    float prod = 1.0f;
    int nsat = 0;
    for (int j = lane; j < nc; j += 32) p24_neg_factor(cls[j], eo1, prod, nsat);
    prod = warp_prod(prod);
    nsat = warp_sum_i(nsat);
    if (!(prod > 1e-30f)) {  // pathological logits: fall back to the term-by-term sum
        const float obj_sig = 1.0f / eo1;
        float s = 0.0f;
        for (int j = lane; j < nc; j += 32) s += p24_bce_neg(p24_joint_prob(cls[j], obj_sig));
        return warp_sum(s);
    }
    return -logf(prod) + 100.0f * (float)nsat;
}

// ... by the 8 lanes of a group
__device__ __noinline__ float group_cls_neg_sum(const float* __restrict__ cls, int nc, float eo1, unsigned m) {
    const int sub = threadIdx.x & 7;
    float prod = 1.0f;
    int nsat = 0;
    for (int j = sub; j < nc; j += 8) p24_neg_factor(cls[j], eo1, prod, nsat);
    prod = group_prod(prod, m);
    nsat = group_sum_i(nsat, m);
    if (!(prod > 1e-30f)) {
        const float obj_sig = 1.0f / eo1;
        float s = 0.0f;
        for (int j = sub; j < nc; j += 8) s += p24_bce_neg(p24_joint_prob(cls[j], obj_sig));
        return group_sum(s, m);
    }
    return -logf(prod) + 100.0f * (float)nsat;
}

// ... by a single thread (rare slow paths)
__device__ __noinline__ float thread_cls_neg_sum(const float* __restrict__ cls, int nc, float eo1) {
    const float obj_sig = 1.0f / eo1;
    float s = 0.0f;
    for (int j = 0; j < nc; ++j) s += p24_bce_neg(p24_joint_prob(cls[j], obj_sig));
    return s;
}

// class cost of one (GT class, anchor) pair from the anchor's all-negative sum
__device__ __forceinline__ float cls_cost_from(float neg_sum, float cls_logit_c, float obj_sig) {
    const float pc = p24_joint_prob(cls_logit_c, obj_sig);
    return (neg_sum - p24_bce_neg(pc)) + p24_bce_pos(pc);
}

// -------------------------------------------------------------------------------------------
// k_pass, anchor CTAs: one CTA per 256-anchor tile
// -------------------------------------------------------------------------------------------
#define ITEM_CAP 704
#define ROW_CH 27  // channels 0..26 of a head row are read here: centre, 24 radii, objectness

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

struct AnchorShared {
    float row[P24_WARPS][ROW_CH][33];
    int cand[P24_THREADS];
    int seedA[P24_SEEDS * 128];
    double red[P24_WARPS];
    float box[P24_WARPS][5];
    int wcnt[P24_WARPS];
    int nitems, n;
};

__device__ __forceinline__ void anchor_part(const Params& p, float4* s_dyn4, AnchorShared& S, int b, int tile, bool first) {
    float* s_gt = reinterpret_cast<float*>(s_dyn4);  // [n * GT_REC_HEAD]: everything but the ray lengths
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int a = tile * P24_THREADS + tid;
    const bool active = a < p.A;
    // the work list lives in rows 5.. of S.row (free between the row reduction and the seed search)
    unsigned* s_items = reinterpret_cast<unsigned*>(&S.row[0][5][0]);  // 726 floats per warp region; ITEM_CAP <= 726

    // ---- the tile's rows: each warp reads its 32 rows, 27 contiguous floats per row (coalesced), straight into
    // shared memory with cp.async; the GT records are computed while they are in flight -------------------------
    const float* img = p.outputs + (long long)b * p.img_stride;
    {
        const int a0 = tile * P24_THREADS + warp * 32;
        const int nrow = min(32, p.A - a0);
        if (lane < ROW_CH) {
            for (int r = 0; r < nrow; ++r) cp_async4(&S.row[warp][lane][r], img + (long long)(a0 + r) * p.row_stride + lane);
        }
    }
    float st = 1.f, xs = 0.f, ys = 0.f;
    if (active) {
        st = p.strides[a];
        xs = p.x_shifts[a];
        ys = p.y_shifts[a];
    }
    if (tid == 0) S.nitems = 0;
    S.cand[tid] = 0;
    TMARK(0, b * p.tiles + tile, 0);
    if (first) pdl_wait();  // the records come from k_gt_prep
    TMARK(0, b * p.tiles + tile, 1);
    const int n = p.num_gt[b];
    {
        const float4* gsrc = reinterpret_cast<const float4*>(p.gt_rec + (long long)b * p.Lmax * GT_REC);
        for (int i = tid; i < n * (GT_REC_HEAD / 4); i += P24_THREADS) {
            const int g = i / (GT_REC_HEAD / 4), q = i - g * (GT_REC_HEAD / 4);
            s_dyn4[i] = gsrc[g * (GT_REC / 4) + q];
        }
    }
    cp_async_wait_all();
    __syncthreads();
    TMARK(0, b * p.tiles + tile, 2);

    float pcx = 0.f, pcy = 0.f, rpmax = 0.f, rpmin = INFINITY, rpsum = 0.f, rp2sum = 0.f, obj = 0.f;
    const float xc = p24_anchor_centre(xs, st);
    const float yc = p24_anchor_centre(ys, st);
    if (active) {
        pcx = S.row[warp][0][lane];
        pcy = S.row[warp][1][lane];
#pragma unroll
        for (int c = 2; c < 26; ++c) {
            const float v = S.row[warp][c][lane];
            rpmax = fmaxf(rpmax, v);
            rpmin = fminf(rpmin, v);
            rpsum += v;
            rp2sum = fmaf(v, v, rp2sum);
        }
        obj = S.row[warp][26][lane];
    }
    double objpart = active ? (double)p24_bce_logits(obj, 0.0f) : 0.0;
    __syncthreads();  // the radii rows of S.row are recycled as the work list from here on

    // ---- one pass over the GTs: centre windows, the inscribed-disc accept, and a bit mask of the GTs whose reject
    // radius the anchor is inside (the only ones that may need a polygon test) ------------------------------------
    bool cheap = false;
    const bool no_prune = (p.flags & P24_F_NO_PRUNE) != 0;
    unsigned near[4] = {0u, 0u, 0u, 0u};  // GTs 0..127; beyond that every GT is tested in place (see below)
    {
        const float r25 = 2.5f * st + 1e-3f * st;  // conservative pre-filter radius of the window test
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            unsigned m = 0u;
            const int ge = min(32, n - w * 32);
            for (int j = 0; j < ge; ++j) {
                const int g = w * 32 + j;
                const float4 h = s_dyn4[g * (GT_REC_HEAD / 4)];
                const float dx = h.x - xc, dy = h.y - yc;
                const float d2 = fmaf(dx, dx, dy * dy);
                cheap |= d2 < h.z;
                m |= (d2 <= h.w ? 1u : 0u) << j;
                if (fmaxf(fabsf(dx), fabsf(dy)) < r25 && p24_in_centre(h.x, h.y, xc, yc, st)) cheap = true;
            }
            near[w] = no_prune ? (ge >= 32 ? 0xFFFFFFFFu : ((ge > 0 ? (1u << ge) : 1u) - 1u)) : m;
        }
        for (int g = 128; g < n; ++g) {  // more than 128 GTs: windows and discs of the rest
            const float4 h = s_dyn4[g * (GT_REC_HEAD / 4)];
            const float dx = h.x - xc, dy = h.y - yc;
            cheap |= fmaf(dx, dx, dy * dy) < h.z;
            if (fmaxf(fabsf(dx), fabsf(dy)) < r25 && p24_in_centre(h.x, h.y, xc, yc, st)) cheap = true;
        }
    }
    cheap = cheap && active;

    // ---- anchors not yet accepted need a polygon test against every near GT; the tests go to a work list so that
    // all threads stay busy (a full list is handled in place) ---------------------------------------------------
    bool mine = false;
    if (active && (!cheap || no_prune)) {
        for (int w = 0; w < 4; ++w) {
            unsigned m = near[w];
            while (m) {
                const int g = w * 32 + __ffs(m) - 1;
                m &= m - 1;
                const int slot = atomicAdd(&S.nitems, 1);
                if (slot < ITEM_CAP) {
                    s_items[slot] = (unsigned)tid | ((unsigned)g << 8);
                } else if (!mine) {
                    const float* rec = s_gt + g * GT_REC_HEAD;
                    mine = no_prune ? p24_in_polygon_exact(rec + GT_VX, rec + GT_VY, xc, yc)
                                    : p24_in_polygon(rec + GT_VX, rec + GT_VY, xc, yc);
                }
            }
        }
        for (int g = 128; g < n && !mine; ++g) {  // more than 128 GTs: test the rest in place
            const float4 h = s_dyn4[g * (GT_REC_HEAD / 4)];
            const float dx = h.x - xc, dy = h.y - yc;
            if (no_prune || fmaf(dx, dx, dy * dy) <= h.w) {
                const float* rec = s_gt + g * GT_REC_HEAD;
                mine = no_prune ? p24_in_polygon_exact(rec + GT_VX, rec + GT_VY, xc, yc)
                                : p24_in_polygon(rec + GT_VX, rec + GT_VY, xc, yc);
            }
        }
    }
    if (mine) S.cand[tid] = 1;
    __syncthreads();
    TMARK(0, b * p.tiles + tile, 3);
    const int nitems = min(S.nitems, ITEM_CAP);
    for (int i = tid; i < nitems; i += P24_THREADS) {
        const unsigned it = s_items[i];
        const int al = it & 0xFF;
        if (((volatile int*)S.cand)[al]) continue;  // already a candidate through another GT
        const int g = it >> 8;
        const float* rec = s_gt + g * GT_REC_HEAD;
        const int aa = tile * P24_THREADS + al;
        const float st2 = p.strides[aa];
        const float axc = p24_anchor_centre(p.x_shifts[aa], st2);
        const float ayc = p24_anchor_centre(p.y_shifts[aa], st2);
        const bool in = no_prune ? p24_in_polygon_exact(rec + GT_VX, rec + GT_VY, axc, ayc)
                                 : p24_in_polygon(rec + GT_VX, rec + GT_VY, axc, ayc);
        if (in) S.cand[al] = 1;
    }
    __syncthreads();

    TMARK(0, b * p.tiles + tile, 4);
    // ---- compacted candidate list of the tile (deterministic order) + per-anchor scratch reset ----------
    const bool cand = active && (n > 0) && (cheap || S.cand[tid]);
    // ---- seeds of the top-10 search: lane g of every warp ranks the warp's candidates for GT g by the first-order
    // proxy q = (mean rg^2 + mean rp^2) / (mean rg + mean rp + d)^2 of the pair value (value ~ 1 - q / 3 for far
    // pairs; the smallest q are almost always the true top-10) and records the largest t = rpmax + d -----------
    S.row[warp][2][lane] = rp2sum * (1.0f / 24.0f);
    S.row[warp][3][lane] = rpsum * (1.0f / 24.0f);
    S.row[warp][4][lane] = cand ? (rpmin < 0.25f ? INFINITY : rpmax) : -1.0f;
    __syncthreads();
    {
        // bounding box of the warp's candidates (GT independent): k_match bounds t = rpmax + d with it
        {
            const float rz = cand ? (rpmin < 0.25f ? INFINITY : rpmax) : P24_NEG_INF;
            const float bx0 = -warp_max(cand ? -pcx : P24_NEG_INF), bx1 = warp_max(cand ? pcx : P24_NEG_INF);
            const float by0 = -warp_max(cand ? -pcy : P24_NEG_INF), by1 = warp_max(cand ? pcy : P24_NEG_INF);
            const float rzm = warp_max(rz);
            if (lane == 0) {
                S.box[warp][0] = bx0;
                S.box[warp][1] = bx1;
                S.box[warp][2] = by0;
                S.box[warp][3] = by1;
                S.box[warp][4] = rzm;
            }
        }
        // warp w ranks a quarter of the tile's anchors, i = 8 j + ((w - j) & 7) for every 4th j: neighbouring anchors
        // (nearly equal proxies) land in different warps, and a quarter sample is enough for seeds (any candidate
        // is a valid seed; better ones only make the bracket tighter).  Per-warp results go to the free rows of S.row.
        float* wres = &S.row[warp][5][0];  // [n][4]: q1, a1, q2, a2   (22 * 33 = 726 floats: n <= 181)
        for (int g = lane; g < n; g += 32) {
            const float* rec = s_gt + g * GT_REC_HEAD;
            const float gcx = rec[GT_CX], gcy = rec[GT_CY], rgms = rec[GT_RGMS], rgmean = rec[GT_RGMEAN];
            float q1 = P24_POS_INF, q2 = P24_POS_INF;
            int a1 = -1, a2 = -1;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = 4 * jj + (warp & 3);
                const int i = 8 * j + ((warp - j) & 7);
                const int wj = i >> 5, lj = i & 31;
                const float rz = S.row[wj][4][lj];
                if (rz < 0.0f) continue;  // not a candidate
                const float dx = gcx - S.row[wj][0][lj], dy = gcy - S.row[wj][1][lj];
                const float d2 = fmaxf(fmaf(dx, dx, dy * dy), 1e-12f);
                const float d = d2 * rsqrtf(d2);
                const float den = (rgmean + S.row[wj][3][lj]) + d;
                const float q = __fdividef(rgms + S.row[wj][2][lj], den * den);
                const int aj = tile * P24_THREADS + i;
                if (q < q1) {
                    q2 = q1;
                    a2 = a1;
                    q1 = q;
                    a1 = aj;
                } else if (q < q2) {
                    q2 = q;
                    a2 = aj;
                }
            }
            if (g < 181) {
                wres[g * 4 + 0] = q1;
                wres[g * 4 + 1] = __int_as_float(a1);
                wres[g * 4 + 2] = q2;
                wres[g * 4 + 3] = __int_as_float(a2);
            }
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, cand);
    if (lane == 0) S.wcnt[warp] = __popc(bal);
    objpart = warp_sum_d(objpart);
    if (lane == 0) S.red[warp] = objpart;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < P24_WARPS; ++w) {
        const int c = S.wcnt[w];
        base += (w < warp) ? c : 0;
        total += c;
    }
    const long long blk = (long long)b * p.tiles + tile;
    // the tile's two best seeds per GT (merge of the 8 warps), evaluated right away (8-lane groups): what
    // k_match brackets the top-10 sum with.  128 GTs at a time.
    for (int g0 = 0; g0 < n; g0 += 128) {
        const int gn = min(128, n - g0);
        if (tid < gn) {
            const int g = g0 + tid;
            float qb[P24_SEEDS];
            int ab[P24_SEEDS];
#pragma unroll
            for (int u = 0; u < P24_SEEDS; ++u) {
                qb[u] = P24_POS_INF;
                ab[u] = -1;
            }
            if (g < 181) {
#pragma unroll
                for (int w = 0; w < P24_WARPS; ++w) {
                    const float* e = &S.row[w][5][0] + g * 4;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        float q = e[2 * u];
                        int aq = __float_as_int(e[2 * u + 1]);
#pragma unroll
                        for (int r = 0; r < P24_SEEDS; ++r) {  // sorted insert
                            if (q < qb[r]) {
                                const float tq = qb[r];
                                const int ta = ab[r];
                                qb[r] = q;
                                ab[r] = aq;
                                q = tq;
                                aq = ta;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < P24_SEEDS; ++u) S.seedA[P24_SEEDS * tid + u] = ab[u];
        }
        __syncthreads();
        {
            const unsigned gm = group_mask();
            const int grp = tid >> 3, sub = tid & 7;
            for (int t = grp; t < P24_SEEDS * gn; t += P24_THREADS / 8) {
                const int sa = S.seedA[t];
                const int g = g0 + t / P24_SEEDS;
                float v = P24_NEG_INF;
                if (sa >= 0)
                    v = group_pair_value_lb(p.gt_rec + ((long long)b * p.Lmax + g) * GT_REC, img + (long long)sa * p.row_stride, gm);
                if (sub == 0)
                    p.sval[((long long)b * p.Lmax + g) * P24_SEEDS * p.tiles + P24_SEEDS * tile + (t % P24_SEEDS)] = v;
            }
        }
        __syncthreads();
    }
    if (tid == 32) {
        float bx0 = INFINITY, bx1 = -INFINITY, by0 = INFINITY, by1 = -INFINITY, rzm = -INFINITY;
        for (int w = 0; w < P24_WARPS; ++w) {
            bx0 = fminf(bx0, S.box[w][0]);
            bx1 = fmaxf(bx1, S.box[w][1]);
            by0 = fminf(by0, S.box[w][2]);
            by1 = fmaxf(by1, S.box[w][3]);
            rzm = fmaxf(rzm, S.box[w][4]);
        }
        float4* dst = reinterpret_cast<float4*>(p.tbox + blk * 8);
        dst[0] = make_float4(bx0, bx1, by0, by1);
        dst[1] = make_float4(rzm, 0.0f, 0.0f, 0.0f);
    }
    if (lane == 0) {
        // the warp's segment of the candidate list: box of its candidates' predicted centres, largest rpmax, extent
        float4* dst = reinterpret_cast<float4*>(p.seg + (blk * P24_WARPS + warp) * 8);
        dst[0] = make_float4(S.box[warp][0], S.box[warp][1], S.box[warp][2], S.box[warp][3]);
        dst[1] = make_float4(S.box[warp][4], __int_as_float(S.wcnt[warp]), __int_as_float(base), __uint_as_float(bal));
    }
    if (cand) {
        const int rank = base + __popc(bal & ((1u << lane) - 1u));
        // a prediction with a tiny radius disables the bound filter for its pairs: rpmax = +inf
        p.clist[blk * P24_THREADS + rank] = make_float4(pcx, pcy, rpmin < 0.25f ? INFINITY : rpmax, __int_as_float(a));
    }
    if (active) {
        // every anchor starts as background; k_resolve_loss overwrites the claimed ones
        const long long o = (long long)b * p.A + a;
        p.claim_cnt[o] = 0;
        p.fg_mask[o] = 0;
        p.matched_gt[o] = -1;
        p.pred_iou[o] = 0.0f;
    }
    if (tid == 0) {
        p.ccount[blk] = total;
        double t = 0.0;
        for (int w = 0; w < P24_WARPS; ++w) t += S.red[w];
        p.obj_part[blk] = t;
    }
    TMARK(0, b * p.tiles + tile, 5);
}

// -------------------------------------------------------------------------------------------
// per-GT helpers of k_match
// -------------------------------------------------------------------------------------------
// GT g selects anchor a: count the claim; the first claimant also puts the anchor on the image's claimed list
__device__ __forceinline__ void claim_anchor(const Params& p, int b, int a, int g) {
    const long long o = (long long)b * p.A + a;
    const int old = atomicAdd(&p.claim_cnt[o], 1);
    p.claim_gt[o] = g;
    if (old == 0) {
        const int slot = atomicAdd(&p.nclaimed[b], 1);
        if (slot < P24_TOPK * p.Lmax) p.claimed[(long long)b * P24_TOPK * p.Lmax + slot] = a;
        else atomicOr(p.err_flag, 2);
    }
}

#define HIT_CAP 2112
#define MATCH_WCAP (25 * P24_MAX_LEVELS)  // at most 5 x 5 cells per level pass the window test
#define MAX_TILES 1024  // candidate counts of an image kept in shared memory (A <= 262144)

// Upper bound of the pair value as a function of t = rpmax + d: any ray has loss <= max(1, 2 - 4 rg^2 / (rg + rp + d)^2)
// (nested rays: loss <= 1; partial and apart rays: loss <= 2 - uni/cs; DESIGN.md "top-10 bracket").
// One thread evaluates it from the GT record in shared memory.
__device__ __forceinline__ float bound_H_thread(const float* __restrict__ rec, float t) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < P24_RAYS; ++k) {
        const float rg = rec[GT_RG + k];
        const float q = rg + t;
        s += fmaxf(1.0f, 2.0f - __fdividef(4.0f * rg * rg, q * q));
    }
    return s * (1.0f / 48.0f);
}

#define EXACT_CAP 2048   // candidates the exact path holds bounds for at a time
#define EXACT_QCAP 512   // warp segments the exact path can keep (more -> brute force)
struct MatchShared {
    float rec[GT_REC];
    int ccount[MAX_TILES];
    int hit[HIT_CAP];       // brute force: exact values; bracket: staged tile boxes; exact path: anchors
    float ev[1024];         // bracket: staged seed values; exact path: exact values of the survivors
    float ub[EXACT_CAP];    // exact path: per-candidate bounds
    float lb[EXACT_CAP];
    int qseg[2 * EXACT_QCAP];  // exact path: kept segments (index, candidate ballot); later the survivors of the threshold
    float qsu[EXACT_QCAP];     // their value bounds, counts, positions in decreasing order of the bound
    int qcnt[EXACT_QCAP];
    int qorder[EXACT_QCAP];
    int hist[32];
    float top[P24_TOPK];
    KV kv[MATCH_WARPS];
    float wmax[MATCH_WARPS];
    int cnt, nhit, nev, k, slow, nvalid, overflow;
    float T, L, tau, tmax, dmax;
    int wanchor[MATCH_WCAP];   // the GT's valid (in window, in polygon) anchors and their costs
    float wcost[MATCH_WCAP];
    int wrank[MATCH_WCAP];     // how many valid pairs cost less
};

template <bool MAX>
__device__ __forceinline__ KV match_block_select(KV x, KV* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    x = warp_select<MAX>(x);
    __syncthreads();
    if (lane == 0) s_red[warp] = x;
    __syncthreads();
    KV y = s_red[lane < MATCH_WARPS ? lane : 0];
    y = warp_select<MAX>(y);
    return y;
}

// the `want` largest of vals[0..n) into S.top (descending) by one warp; destroys vals; returns the count
__device__ int warp_select_top(float* vals, int n, int want, float* top) {
    const int lane = threadIdx.x & 31;
    int got = 0;
    for (int r = 0; r < want; ++r) {
        KV best = {P24_NEG_INF, 0x7fffffff};
        for (int i = lane; i < n; i += 32) {
            const float v = vals[i];
            if (kv_gt(v, i, best.v, best.i)) {
                best.v = v;
                best.i = i;
            }
        }
        best = warp_select<true>(best);
        if (best.i == 0x7fffffff) break;
        if (lane == 0) {
            top[r] = best.v;
            vals[best.i] = P24_NEG_INF;
        }
        __syncwarp();
        ++got;
    }
    return got;
}

// Brute force (no usable threshold, P24_F_NO_FILTER, or list overflow): exact value of EVERY candidate, in chunks,
// carrying the best kc forward.
__device__ __noinline__ float topk_sum_bruteforce(const Params& p, MatchShared& S, int b, int kc) {
    const int tid = threadIdx.x, warp = tid >> 5;
    const unsigned gm = group_mask();
    const int grp = tid >> 3, sub = tid & 7;
    const float* img = p.outputs + (long long)b * p.img_stride;
    float* vals = reinterpret_cast<float*>(S.hit);  // HIT_CAP floats
    int ncarry = 0;
    const int tiles_per_chunk = (HIT_CAP - P24_TOPK) / P24_THREADS;
    for (int t0 = 0; t0 < p.tiles; t0 += tiles_per_chunk) {
        const int t1 = min(t0 + tiles_per_chunk, p.tiles);
        __syncthreads();
        if (tid < ncarry) vals[tid] = S.top[tid];
        if (tid == 0) S.nhit = ncarry;
        __syncthreads();
        for (int tl = t0; tl < t1; ++tl) {
            const long long blk = (long long)b * p.tiles + tl;
            const int c = S.ccount[tl];
            for (int i = grp; i < c; i += MATCH_GROUPS) {
                const int a = __float_as_int(p.clist[blk * P24_THREADS + i].w);
                const float v = group_pair_value(S.rec, img + (long long)a * p.row_stride, gm);
                if (sub == 0) vals[atomicAdd(&S.nhit, 1)] = (v == v) ? v : P24_POS_INF;  // NaN sorts first (torch.topk)
            }
        }
        __syncthreads();
        if (warp == 0) warp_select_top(vals, S.nhit, kc, S.top);
        __syncthreads();
        ncarry = min(kc, S.nhit);
    }
    float ksum = 0.0f;
    for (int i = 0; i < ncarry; ++i) ksum = ksum + (S.top[i] == P24_POS_INF ? NAN : S.top[i]);
    return ksum;
}

// Upper bound of the pair value as a function of the centre distance d alone.  For an apart ray the loss is
// 2 - 4 (rg^2 + rp^2) / (rg + rp + d)^2 (and that expression bounds partial rays too, see bound_H_thread); over all
// rp > 0 the fraction is smallest at rp* = rg^2 / (rg + d), where it equals rg^2 / ((rg + d)^2 + rg^2).  So every
// ray has loss <= max(1, 2 - 4 rg^2 / ((rg + d)^2 + rg^2)) whatever the prediction: monotone in d, and tight for large
// GTs, where bound_H_thread (which pays for the largest predicted radius of a whole tile) is loose.
__device__ __forceinline__ float bound_Hstar_thread(const float* __restrict__ rec, float d) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < P24_RAYS; ++k) {
        const float rg = rec[GT_RG + k];
        const float q = rg + d;
        s += fmaxf(1.0f, 2.0f - __fdividef(4.0f * rg * rg, fmaf(q, q, rg * rg)));
    }
    return s * (1.0f / 48.0f);
}

// Exact top-10 sum when the bracket is not conclusive (about 1 GT in 100; the large ones, whose values sit near 0.9
// and whose sums sit near 9), in one pass over what can matter:
//  1. every warp segment of the candidate lists (32 anchors: a short run of one grid row) whose value bound
//     min(H(t), H*(d)) over its box reaches T (the 10th best seed value, a certified lower bound of the 10th largest
//     value) is kept: the far corners of the image as seen from the GT;
//  2. their candidates get per-candidate bounds (one thread each: ub = the apart formula, which bounds every ray;
//     for a pair whose rays are all apart -- the reference's own fp32 comparison -- the value is ub to within 3e-6,
//     so ub - 5e-5 is a certified lower bound);
//  3. the 10th largest lower bound (two rounds of a 32-bin histogram) replaces T;
//  4. the candidates whose ub still reaches it (a handful) are evaluated exactly (8-lane groups);
//  5. the 10 largest exact values are summed in descending order (torch.topk order).
// Returns NaN when the lists do not fit (the caller falls back to brute force).
__device__ __noinline__ float topk_sum_exact(const Params& p, MatchShared& S, int b) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned gm = group_mask();
    const int grp = tid >> 3, sub = tid & 7;
    const float gcx = S.rec[GT_CX], gcy = S.rec[GT_CY];
    const float* img = p.outputs + (long long)b * p.img_stride;
    const float T0 = S.T;
    if (tid == 0) {
        S.nhit = 0;  // qualifying segments
        S.cnt = 0;   // their candidates
        S.nev = 0;
        S.overflow = 0;
    }
    __syncthreads();
    {
        const float4* sg = reinterpret_cast<const float4*>(p.seg + (long long)b * p.tiles * P24_WARPS * 8);
        const int nseg = p.tiles * P24_WARPS;
        for (int si = tid; si < nseg; si += MATCH_THREADS) {
            const float4 bx = sg[2 * si];
            const float4 r1 = sg[2 * si + 1];
            const int cnt = __float_as_int(r1.y);
            if (cnt <= 0) continue;
            const float fx = fmaxf(fabsf(gcx - bx.x), fabsf(gcx - bx.y));
            const float fy = fmaxf(fabsf(gcy - bx.z), fabsf(gcy - bx.w));
            const float dm = sqrtf(fmaf(fx, fx, fy * fy)) * 1.0001f + 0.01f;
            const float u = fminf(bound_H_thread(S.rec, r1.x + dm), bound_Hstar_thread(S.rec, dm)) + 2e-5f;
            if (u < T0 && r1.x < 60000.0f) continue;  // (NaN bounds and tiny predicted radii stay in)
            const int q = atomicAdd(&S.nhit, 1);
            const int pos = atomicAdd(&S.cnt, cnt);
            if (q < EXACT_QCAP) {
                S.qseg[2 * q] = si;
                S.qseg[2 * q + 1] = __float_as_int(r1.w);  // which of the segment's 32 anchors are candidates
                S.qsu[q] = (u == u && r1.x < 60000.0f) ? u : P24_POS_INF;
                S.qcnt[q] = cnt;
                S.qorder[q] = pos;  // the segment's first slot when everything fits (the usual case)
            } else {
                S.overflow = 1;
            }
        }
    }
    __syncthreads();
    if (S.overflow) return NAN;
    const int nq = S.nhit;
    const bool all_fit = S.cnt <= EXACT_CAP;
    __syncthreads();
    if (all_fit) {
        for (int q = tid; q < nq; q += MATCH_THREADS) {
            S.qcnt[q] = S.qorder[q];
            S.qorder[q] = q;
        }
        if (tid == 0) {
            S.k = nq;
            S.L = P24_NEG_INF;
        }
    }
    // otherwise: the kept segments in decreasing order of their bound; the first ones that fit EXACT_CAP candidates are
    // examined, and the best bound among the others (S.L) must end up below the refined threshold
    for (int q = tid; q < nq && !all_fit; q += MATCH_THREADS) {
        const float uq = S.qsu[q];
        int rank = 0;
        for (int j = 0; j < nq; ++j) rank += kv_gt(S.qsu[j], j, uq, q) ? 1 : 0;
        S.qorder[rank] = q;
    }
    __syncthreads();
    if (warp == 0 && !all_fit) {
        int run = 0, ncut = nq;  // candidates so far; number of examined segments
        for (int r0 = 0; r0 < nq && ncut == nq; r0 += 32) {
            const int r = r0 + lane;
            const int q = r < nq ? S.qorder[r] : 0;
            const int c = r < nq ? S.qcnt[q] : 0;
            int inc = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, off);
                if (lane >= off) inc += t;
            }
            const bool fits = r < nq && run + inc <= EXACT_CAP;
            if (fits) S.qcnt[q] = run + inc - c;  // from here on: the segment's first slot
            const unsigned fm = __ballot_sync(0xffffffffu, fits);
            const unsigned vm = __ballot_sync(0xffffffffu, r < nq);
            const int nfit = __popc(fm);  // fits is a prefix of the valid lanes (counts are positive)
            if (fm != vm) ncut = r0 + nfit;
            const int last = __shfl_sync(0xffffffffu, inc, nfit > 0 ? nfit - 1 : 0);
            if (nfit > 0) run += last;
        }
        if (lane == 0) {
            S.k = ncut;
            S.cnt = run;
            S.L = ncut < nq ? S.qsu[S.qorder[ncut]] : P24_NEG_INF;
        }
    }
    __syncthreads();
#ifdef P24_TIMING
    if (tid == 0) {
        g_tstamp[1][b * 20 + (int)blockIdx.y][9] = S.nhit;
        g_tstamp[1][b * 20 + (int)blockIdx.y][10] = S.cnt;
        g_tstamp[1][b * 20 + (int)blockIdx.y][11] = S.k;
    }
    TMARK(1, b * 20 + (int)blockIdx.y, 16);
#endif
    const int n1 = S.cnt, ncut = S.k;
    const float su_rest = S.L;
    // the anchors of the examined segments (a segment is 32 consecutive anchors; its candidates are the set bits)
    for (int r = warp; r < ncut; r += MATCH_WARPS) {
        const int q = S.qorder[r];
        const int si = S.qseg[2 * q], pos = S.qcnt[q];
        const unsigned m = (unsigned)S.qseg[2 * q + 1];
        if ((m >> lane) & 1u) S.hit[pos + __popc(m & ((1u << lane) - 1u))] = si * 32 + lane;
    }
    __syncthreads();
    // per-candidate bounds, one candidate per thread and pass: all loads of a pass are independent
    for (int i = tid; i < n1; i += MATCH_THREADS) {
        const int a = S.hit[i];
        const float* row = img + (long long)a * p.row_stride;
        float rpv[P24_RAYS];
#pragma unroll
        for (int k = 0; k < P24_RAYS; ++k) rpv[k] = row[2 + k];  // one round trip for the whole row
        const float d = p24_centre_dist(gcx, gcy, row[0], row[1]);
        float u = 0.0f, rpmin = INFINITY;
        bool apart = true;
#pragma unroll
        for (int k = 0; k < P24_RAYS; ++k) {
            const float rg = S.rec[GT_RG + k], rp = rpv[k];
            u += p24_ray_loss_ub(rg, rp, d);
            apart = apart && (d >= rg + rp);
            rpmin = fminf(rpmin, rp);
        }
        u = u * (1.0f / 48.0f) + 2e-5f;
        const bool trust = rpmin >= 0.25f && u == u;  // tiny predicted radii / NaN: evaluated exactly, no bounds
        S.ub[i] = trust ? u : P24_POS_INF;
        S.lb[i] = (trust && apart) ? u - 5e-5f : P24_NEG_INF;
    }
    __syncthreads();
    TMARK(1, b * 20 + (int)blockIdx.y, 17);
    // the 10th largest lower bound to within (1 - T0) / 1024: two rounds of a 32-bin histogram, starting from [T0, 1]
    float lo = T0, width = fmaxf(1.0f - T0, 1e-6f);
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        if (tid < 32) S.hist[tid] = 0;
        __syncthreads();
        const float scale = 32.0f / width;
        for (int i = tid; i < n1; i += MATCH_THREADS) {
            const float l = S.lb[i];
            if (l >= lo) atomicAdd(&S.hist[min((int)((l - lo) * scale), 31)], 1);
        }
        __syncthreads();
        // the highest bin whose count, together with the bins above it, reaches 10
        int above = 0, bin = -1;
        for (int j = 31; j >= 0; --j) {
            above += S.hist[j];
            if (above >= P24_TOPK) {
                bin = j;
                break;
            }
        }
        __syncthreads();
        if (bin < 0) break;  // fewer than 10 lower bounds reach lo: lo stays (it is certified by the seeds or the last round)
        lo = lo + (float)bin * (width * (1.0f / 32.0f));
        width = width * (1.0f / 32.0f);
    }
    const float tcur = fmaxf(T0, lo - 1e-6f);
    if (!(su_rest < tcur)) return NAN;  // a segment that was not examined could still hold a top-10 value (rare)
    for (int i = tid; i < n1; i += MATCH_THREADS)
        if (!(S.ub[i] < tcur)) {
            const int at = atomicAdd(&S.nev, 1);
            if (at < 1024) S.qseg[at] = S.hit[i];  // (the segment list is no longer needed)
            else S.overflow = 1;
        }
    __syncthreads();
    TMARK(1, b * 20 + (int)blockIdx.y, 18);
#ifdef P24_TIMING
    if (tid == 0) g_tstamp[1][b * 20 + (int)blockIdx.y][7] = S.nev;
#endif
    if (S.overflow) return NAN;
    const int nsurv = S.nev;
    for (int i0 = 0; i0 < nsurv; i0 += MATCH_GROUPS) {
        const int i = i0 + grp;
        if (i >= nsurv) continue;
        const float v = group_pair_value(S.rec, img + (long long)S.qseg[i] * p.row_stride, gm);
        if (sub == 0) S.ev[i] = (v == v) ? v : P24_POS_INF;  // NaN sorts first (torch.topk)
    }
    __syncthreads();
    TMARK(1, b * 20 + (int)blockIdx.y, 19);
    if (nsurv < P24_TOPK) return NAN;
    for (int i = tid; i < nsurv; i += MATCH_THREADS) {
        const float vi = S.ev[i];
        int rank = 0;
        for (int j = 0; j < nsurv; ++j) rank += kv_gt(S.ev[j], j, vi, i) ? 1 : 0;
        if (rank < P24_TOPK) S.top[rank] = vi;
    }
    __syncthreads();
    float ksum = 0.0f;
    for (int i = 0; i < P24_TOPK; ++i) ksum = ksum + (S.top[i] == P24_POS_INF ? NAN : S.top[i]);
    return ksum;
}

// Spill path (rare: GT with fewer valid anchors than its dynamic k): take `need` more anchors with the
// smallest PENALISED cost among the candidates that are not valid for this GT.  Ties -> lower anchor index.
__device__ __noinline__ void spill_claims(const Params& p, MatchShared& S, int b, int g, int nwin, int need) {
    float lv[P24_TOPK];
    int li[P24_TOPK];
#pragma unroll
    for (int i = 0; i < P24_TOPK; ++i) {
        lv[i] = P24_POS_INF;
        li[i] = 0x7fffffff;
    }
    const int tid = threadIdx.x;
    const float* img = p.outputs + (long long)b * p.img_stride;
    const int c = gt_class(S.rec, p.nc);
    for (int tl = 0; tl < p.tiles; ++tl) {
        const long long blk = (long long)b * p.tiles + tl;
        const int cc = S.ccount[tl];
        for (int i = tid; i < cc; i += MATCH_THREADS) {
            const int a = __float_as_int(p.clist[blk * P24_THREADS + i].w);
            bool isvalid = false;
            for (int j = 0; j < nwin; ++j) isvalid |= (S.wanchor[j] == a && S.wcost[j] != P24_POS_INF);
            if (isvalid) continue;
            const float* row = img + (long long)a * p.row_stride;
            const float eo1 = 1.0f + expf(-row[26]);
            const float neg = thread_cls_neg_sum(row + 27, p.nc, eo1);
            const float v = pair_value_row(S.rec, row);
            const float cost = p24_cost(cls_cost_from(neg, row[27 + c], 1.0f / eo1), v, false);
            if (kv_lt(cost, a, lv[P24_TOPK - 1], li[P24_TOPK - 1])) {
                float cv = cost;
                int ci = a;
#pragma unroll
                for (int q = 0; q < P24_TOPK; ++q) {
                    if (kv_lt(cv, ci, lv[q], li[q])) {
                        const float tv = lv[q];
                        const int ti = li[q];
                        lv[q] = cv;
                        li[q] = ci;
                        cv = tv;
                        ci = ti;
                    }
                }
            }
        }
    }
    for (int r = 0; r < need; ++r) {
        const KV head = {lv[0], li[0]};
        const KV win = match_block_select<false>(head, S.kv);
        if (win.i == 0x7fffffff) break;  // fewer candidates than needed
        if (li[0] == win.i && lv[0] == win.v) {
            claim_anchor(p, b, win.i, g);
#pragma unroll
            for (int q = 0; q < P24_TOPK - 1; ++q) {
                lv[q] = lv[q + 1];
                li[q] = li[q + 1];
            }
            lv[P24_TOPK - 1] = P24_POS_INF;
            li[P24_TOPK - 1] = 0x7fffffff;
        }
    }
}

// -------------------------------------------------------------------------------------------
// k_pass, centre-window CTAs: every (GT, centre-window anchor) pair as an independent 8-lane group task:
// polygon test (inscribed-disc accept, else the reference-order edge terms, 3 per lane), exact pair value and
// SimOTA cost when inside.  The window of a GT is enumerated straight from the level grids (a 7 x 7 block of cells
// per level around the centre holds every anchor that can pass the strict test of losses.py:523-542, which is then
// applied in the reference's own arithmetic), so this part needs nothing from the anchor CTAs and runs beside them.
// Costs land in the GT's window table by slot (level, row, column); the anchor of a slot and the slot of an anchor
// are both computable, which is what k_match (selection) and k_resolve_loss (conflict argmin) rely on.
// -------------------------------------------------------------------------------------------
struct WindowShared {
    float rec[GT_REC];
    int list[P24_WSLOTS];  // slots of the level that pass the centre-window test
    int org[2];
    int npair;
};

// first cell of the 7-wide block that contains every cell centre within 2.5 strides of c (one cell of slack per side)
__device__ __forceinline__ int window_origin(float c, float st) {
    float v = floorf(c / st) - 3.0f;
    v = fminf(fmaxf(v, -1.0e6f), 1.0e6f);  // NaN -> -1e6: an empty window
    return (int)v;
}

// one work item: the window of GT g of image b on level l
__device__ __forceinline__ void window_part(const Params& p, WindowShared& S, int b, int g, int l) {
    const int tid = threadIdx.x;
    const unsigned gm = group_mask();
    const int grp = tid >> 3, sub = tid & 7;
    const float* img = p.outputs + (long long)b * p.img_stride;
    const Level lv = p.lev[l];
    if (tid < GT_REC) S.rec[tid] = p.gt_rec[((long long)b * p.Lmax + g) * GT_REC + tid];
    if (tid == GT_REC) S.npair = 0;
    __syncthreads();
    const float gcx = S.rec[GT_CX], gcy = S.rec[GT_CY];
    float* tab = p.wtab + ((long long)b * p.Lmax + g) * P24_WT_STRIDE;
    const float lst = p.strides[lv.off];
    const int ox = window_origin(gcx, lst), oy = window_origin(gcy, lst);
    if (tid < 2) tab[P24_WT_HDR + 2 * l + tid] = __int_as_float(tid ? oy : ox);
    if (tid < P24_WSLOTS) {
        const int sy = tid / P24_WSIDE, sx = tid - sy * P24_WSIDE;
        const int ix = ox + sx, iy = oy + sy;
        bool in = false;
        if (ix >= 0 && ix < lv.W && iy >= 0 && iy < lv.H) {
            const int a = lv.off + iy * lv.W + ix;
            const float st = p.strides[a];
            in = p24_in_centre(gcx, gcy, p24_anchor_centre(p.x_shifts[a], st), p24_anchor_centre(p.y_shifts[a], st), st);
        }
        tab[l * P24_WSLOTS + tid] = P24_POS_INF;
        if (in) S.list[atomicAdd(&S.npair, 1)] = tid;
    }
    __syncthreads();
    const int npair = S.npair;
    const int c = gt_class(S.rec, p.nc);
    for (int wi = grp; wi < npair; wi += P24_THREADS / 8) {
        const int t = S.list[wi];
        const int sy = t / P24_WSIDE, sx = t - sy * P24_WSIDE;
        const int a = lv.off + (oy + sy) * lv.W + (ox + sx);
        const float* row = img + (long long)a * p.row_stride;
        const float st = p.strides[a];
        const float xs = p.x_shifts[a], ys = p.y_shifts[a];
        const float pcx = row[0], pcy = row[1], obj = row[26], clsc = row[27 + c];
        float rp[3], cl[10];
#pragma unroll
        for (int q = 0; q < 3; ++q) rp[q] = row[2 + sub * 3 + q];
#pragma unroll
        for (int q = 0; q < 10; ++q) cl[q] = (sub + 8 * q < p.nc) ? row[27 + sub + 8 * q] : 0.0f;
        const float xc = p24_anchor_centre(xs, st);
        const float yc = p24_anchor_centre(ys, st);
        bool inside = true;
        {
            // inside the inscribed disc the angle sum is >= 360 (see warp_gt_record): no edge terms needed
            const float ddx = gcx - xc, ddy = gcy - yc;
            if (!(fmaf(ddx, ddx, ddy * ddy) < S.rec[GT_RIN2]) || (p.flags & P24_F_NO_PRUNE)) {
                float ang = 0.0f;
#pragma unroll 1
                for (int q = 0; q < 3; ++q) {
                    const int k = sub * 3 + q;
                    const int k2 = (k == P24_RAYS - 1) ? 0 : k + 1;
                    ang = ang + edge_angle(S.rec[GT_VX + k] - xc, S.rec[GT_VY + k] - yc, S.rec[GT_VX + k2] - xc,
                                           S.rec[GT_VY + k2] - yc);
                }
                ang = group_sum(ang, gm);
                inside = ang >= 350.0f;  // losses.py:588
            }
        }
        if (inside) {
            const float d = p24_centre_dist(gcx, gcy, pcx, pcy);
            float sm = 0.0f;
#pragma unroll 1
            for (int q = 0; q < 3; ++q) sm = sm + ray_loss(S.rec[GT_RG + sub * 3 + q], rp[q], d);
            sm = group_sum(sm, gm);
            const float v = (sm / 24.0f) / 2.0f;
            const float eo1 = 1.0f + expf(-obj);
            float neg;
            if (p.nc <= 80) {
                float prod = 1.0f;
                int nsat = 0;
#pragma unroll
                for (int q = 0; q < 10; ++q)
                    if (sub + 8 * q < p.nc) p24_neg_factor(cl[q], eo1, prod, nsat);
                prod = group_prod(prod, gm);
                nsat = group_sum_i(nsat, gm);
                neg = (prod > 1e-30f) ? (-logf(prod) + 100.0f * (float)nsat) : group_cls_neg_sum(row + 27, p.nc, eo1, gm);
            } else {
                neg = group_cls_neg_sum(row + 27, p.nc, eo1, gm);
            }
            float cost = p24_cost(cls_cost_from(neg, clsc, 1.0f / eo1), v, true);
            if (!(cost < 3.0e38f)) cost = 3.0e38f;  // NaN / inf inputs: keep the pair selectable, last
            if (sub == 0) tab[l * P24_WSLOTS + t] = cost;
        }
    }
}

// k_pass: persistent CTAs (one wave) drawing work items from a ticket counter: first the anchor tiles (the long items),
// then the (GT, level) centre-window items.  Launched as a programmatic dependent of k_gt_prep: the first item's rows
// are in flight before the CTA waits for the records.
union PassShared {
    AnchorShared a;
    WindowShared w;
};

__global__ void __launch_bounds__(P24_THREADS, 4) k_pass(const __grid_constant__ Params p) {
    extern __shared__ float4 s_dyn4[];
    __shared__ PassShared S;
    __shared__ int s_item;
    const int n_anchor = p.B * p.tiles;
    bool first = true;
    int leff = 0;  // the batch's largest num_gt (known once k_gt_prep is complete)
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = (int)atomicAdd(&p.ticket[1 + p.B], 1u);
        __syncthreads();
        const int item = s_item;
        if (item < n_anchor) {
            // image-major order keeps an image's tiles (and its records) together in time
            anchor_part(p, s_dyn4, S.a, item / p.tiles, item % p.tiles, first);
            first = false;
        } else {
            if (first) {
                pdl_wait();
                first = false;
            }
            if (leff == 0) leff = (int)__ldcg(&p.ticket[2 + p.B]);
            const int wi = item - n_anchor;
            if (wi >= p.B * leff * p.nlev) break;
            const int l = wi % p.nlev, bg = wi / p.nlev;
            const int b = bg / leff, g = bg - b * leff;
            if (g < p.num_gt[b]) {
                TMARK(4, wi, 0);
                window_part(p, S.w, b, g, l);
                TMARK(4, wi, 1);
            }
        }
    }
    pdl_trigger();
}

// dynamic k is known: record it and claim the k smallest costs among the GT's valid pairs (losses.py:460-464; the valid
// pairs and their ranks -- ties -> lower anchor index -- were prepared while the bracket was being evaluated); spill into
// the penalised regime when there are fewer valid pairs than k.
__device__ __forceinline__ void claim_selected(const Params& p, MatchShared& S, int b, int g, int k) {
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid == 0) p.dyn_k[b * p.Lmax + g] = k;
    const int nv = min(S.nvalid, MATCH_WCAP);
    const int take = min(k, nv);
    if (tid < nv && S.wrank[tid] < take) claim_anchor(p, b, S.wanchor[tid], g);
    if (k > nv) spill_claims(p, S, b, g, nv, k - nv);
}

// -------------------------------------------------------------------------------------------
// k_match: per GT, dynamic k = clamp(int(sum of the 10 largest pair values over the candidates), 1) from the seed
// values of the anchor pass (top-10 bracket; exact filtered / brute-force paths when it is not conclusive), then
// the k smallest costs of its valid pairs -> claims (rank counting; spill into the penalised regime when the GT
// has fewer valid anchors than k)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MATCH_THREADS, 3) k_match(const __grid_constant__ Params p) {
    // images vary fastest over the grid: the low GT rows (the real ones: valid rows come first) are dispatched before
    // the rows beyond num_gt, which leave at once
    const int g = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
#define MCTA (b * 20 + g)
    if (g < 20) TMARK(1, MCTA, 0);
    pdl_trigger();  // k_resolve_loss may become resident
    pdl_wait();
    const int n = p.num_gt[b];
    if (g >= n) {
        if (tid == 0) p.dyn_k[b * p.Lmax + g] = 0;
        return;
    }
    TMARK(1, MCTA, 1);
    __shared__ MatchShared S;
    const int wslot = b * p.Lmax + g;
    if (tid < GT_REC) S.rec[tid] = p.gt_rec[(long long)wslot * GT_REC + tid];
    int cnt = 0;
    for (int tl = tid; tl < p.tiles; tl += MATCH_THREADS) {
        const int c = p.ccount[(long long)b * p.tiles + tl];
        S.ccount[tl] = c;
        cnt += c;
    }
    // everything the bracket reads, staged by all threads in one round trip (S.hit / S.ev are free until a slow path)
    float4* s_tb = reinterpret_cast<float4*>(S.hit);   // [2 * tiles] float4   (tiles <= 256 here, else read in place)
    float* s_sv = S.ev;                                // [P24_SEEDS * tiles]
    const bool staged = p.tiles <= 256;
    if (staged) {
        const float4* tbg = reinterpret_cast<const float4*>(p.tbox + (long long)b * p.tiles * 8);
        for (int i = tid; i < 2 * p.tiles; i += MATCH_THREADS) s_tb[i] = tbg[i];
        const float* svg = p.sval + (long long)wslot * (P24_SEEDS * p.tiles);
        for (int i = tid; i < P24_SEEDS * p.tiles; i += MATCH_THREADS) s_sv[i] = svg[i];
    }
    // the GT's window table (costs of its valid pairs), requested in the same round trip: the selection is prepared by
    // warps 1..7 while warp 0 evaluates the bracket
    const float* tab = p.wtab + (long long)wslot * P24_WT_STRIDE;
    const int nslot = P24_WSLOTS * p.nlev;  // <= 196 < MATCH_THREADS
    float wc = P24_POS_INF, wox = 0.0f, woy = 0.0f;
    if (tid < nslot) {
        const int l = tid / P24_WSLOTS;
        wc = tab[tid];
        wox = tab[P24_WT_HDR + 2 * l];
        woy = tab[P24_WT_HDR + 2 * l + 1];
    }
    if (tid == 0) {
        S.cnt = 0;
        S.nvalid = 0;
    }
    __syncthreads();
    cnt = warp_sum_i(cnt);
    if (lane == 0 && cnt) atomicAdd(&S.cnt, cnt);
    if (wc < P24_POS_INF) {
        const int l = tid / P24_WSLOTS, r = tid - l * P24_WSLOTS;
        const int sy = r / P24_WSIDE, sx = r - sy * P24_WSIDE;
        const int ix = __float_as_int(wox) + sx, iy = __float_as_int(woy) + sy;
        const int slot = atomicAdd(&S.nvalid, 1);
        if (slot < MATCH_WCAP) {
            S.wanchor[slot] = p.lev[l].off + iy * p.lev[l].W + ix;
            S.wcost[slot] = wc;
        } else {
            atomicOr(p.err_flag, 1);
        }
    }
    const float gcx = S.rec[GT_CX], gcy = S.rec[GT_CY];
    __syncthreads();
    TMARK(1, MCTA, 2);
    const int ncand = S.cnt;
    const int kc = min(P24_TOPK, ncand);  // losses.py:452

    // ---- bracket the top-10 sum: L = sum of the 10 best seed values <= S <= 10 * H(t_max) = U (warp 0) ------------
    if (warp == 0) {
        // largest t = rpmax + d over the candidates, bounded per tile by its box
        float tm = P24_NEG_INF, dm = P24_NEG_INF;  // ... and the largest centre distance d
        const float4* tb = staged ? s_tb : reinterpret_cast<const float4*>(p.tbox + (long long)b * p.tiles * 8);
        for (int i = lane; i < p.tiles; i += 32) {
            const float4 bx = tb[2 * i];
            const float rzm = tb[2 * i + 1].x;
            if (rzm > P24_NEG_INF) {
                const float fx = fmaxf(fabsf(gcx - bx.x), fabsf(gcx - bx.y));
                const float fy = fmaxf(fabsf(gcy - bx.z), fabsf(gcy - bx.w));
                const float dd = sqrtf(fmaf(fx, fx, fy * fy)) * 1.00001f;
                tm = fmaxf(tm, rzm + dd);
                dm = fmaxf(dm, dd);
            }
        }
        tm = warp_max(tm);
        dm = warp_max(dm);
        // the 10 largest seed values, summed in descending order (like the reference sums torch.topk's output)
        const int nseed = P24_SEEDS * p.tiles;
        const float* sv = staged ? s_sv : p.sval + (long long)wslot * nseed;
        float v0 = P24_NEG_INF, v1 = P24_NEG_INF, v2 = P24_NEG_INF, v3 = P24_NEG_INF;  // the lane's 4 best seeds
        for (int i = lane; i < nseed; i += 32) {
            float v = sv[i];
            if (v > v0) { const float t0 = v0; v0 = v; v = t0; }
            if (v > v1) { const float t1 = v1; v1 = v; v = t1; }
            if (v > v2) { const float t2 = v2; v2 = v; v = t2; }
            if (v > v3) v3 = v;
        }
        float T = P24_NEG_INF, L = 0.0f;
        int got = 0;
        if (kc == P24_TOPK) {
#pragma unroll 1
            for (int r = 0; r < P24_TOPK; ++r) {
                const float m = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
                const KV best = warp_select<true>(KV{m, lane});
                if (!(best.v > P24_NEG_INF)) break;
                if (lane == best.i) {  // remove one copy of the winner
                    if (v0 == m) v0 = P24_NEG_INF;
                    else if (v1 == m) v1 = P24_NEG_INF;
                    else if (v2 == m) v2 = P24_NEG_INF;
                    else v3 = P24_NEG_INF;
                }
                L = L + best.v;
                T = best.v;
                ++got;
            }
            if (got < P24_TOPK) T = P24_NEG_INF;
        }
        int slow = 1, k = 0;
        const bool usable = T > P24_NEG_INF && !(p.flags & P24_F_NO_FILTER) && S.rec[GT_RGMIN] >= 0.25f && tm < 60000.0f;
        if (usable) {
            // two monotone bounds of any candidate's value: H(t_max) (bound_H_thread) and H*(d_max) (bound_Hstar_thread)
            float term = 0.0f, term2 = 0.0f;
            if (lane < P24_RAYS) {
                const float rg = S.rec[GT_RG + lane];
                const float q = rg + (tm * 1.0001f + 0.01f);
                term = fmaxf(1.0f, 2.0f - __fdividef(4.0f * rg * rg, q * q));
                const float q2 = rg + (dm * 1.0001f + 0.01f);
                term2 = fmaxf(1.0f, 2.0f - __fdividef(4.0f * rg * rg, fmaf(q2, q2, rg * rg)));
            }
            const float U = 10.0f * (fminf(warp_sum(term), warp_sum(term2)) * (1.0f / 48.0f) + 2e-5f);
            const float fl = floorf(L - 1e-4f), fu = floorf(U + 1e-4f);
#ifdef P24_TIMING
            if (lane == 0) {
                g_tstamp[1][MCTA][12] = __float_as_uint(L);
                g_tstamp[1][MCTA][13] = __float_as_uint(U);
                g_tstamp[1][MCTA][14] = __float_as_uint(tm);
                g_tstamp[1][MCTA][15] = __float_as_uint(S.rec[GT_RGMAX]);
            }
#endif
            if (fl == fu && fl >= 1.0f) {
                slow = 0;
                k = (int)fl;
            }
        }
        if (lane == 0) {
            S.slow = slow ? (usable ? 1 : 2) : 0;  // 1: filtered exact path, 2: brute force
            S.k = k;
            S.T = T;
            S.tmax = tm;
        }
    } else {
        // ---- meanwhile: rank of every valid pair by (cost, anchor) -----------------------------------------------------
        const int nv = min(S.nvalid, MATCH_WCAP);
        for (int e = tid - 32; e < nv; e += MATCH_THREADS - 32) {
            const float ci = S.wcost[e];
            const int ai = S.wanchor[e];
            int before = 0;
            for (int j = 0; j < nv; ++j) before += kv_lt(S.wcost[j], S.wanchor[j], ci, ai) ? 1 : 0;
            S.wrank[e] = before;
        }
    }
    __syncthreads();
    TMARK(1, MCTA, 4);
#ifdef P24_TIMING
    if (tid == 0) {
        g_tstamp[1][MCTA][8] = S.slow;
        g_tstamp[1][MCTA][9] = 0;
    }
#endif
    int k;
    if (S.slow) {
        float ksum = NAN;
        if (S.slow == 1) ksum = topk_sum_exact(p, S, b);
        if (!(ksum == ksum)) ksum = topk_sum_bruteforce(p, S, b, kc);
        k = (int)ksum;  // dynamic k = clamp(int(sum of the top-kc values), min=1)   losses.py:454-456
        if (k < 1) k = 1;
    } else {
        k = S.k;
    }
    TMARK(1, MCTA, 5);
    claim_selected(p, S, b, g, min(k, ncand));  // torch.topk would raise beyond the candidate count; clamp instead
    TMARK(1, MCTA, 6);
}

// -------------------------------------------------------------------------------------------
// k_resolve_loss
// -------------------------------------------------------------------------------------------
// normalisation + stateful re-weighting, losses.py:280-345; executed by one warp
__device__ void finalize_warp(const float* sums28, float* state26, float* result54, float* weights_n27) {
    const int lane = threadIdx.x & 31;
    const float nfg = fmaxf(sums28[26], 1.0f);
    const float ngt = fmaxf(sums28[27], 1.0f);
    float loss = 0.0f, e = 0.0f;
    if (lane < 26) {
        loss = sums28[lane] / nfg;  // loss_iou[k], loss_obj, loss_cls
        float r = loss / (state26[lane] + 1e-8f);
        r = fminf(fmaxf(r, 0.0f), 2.0f);
        e = expf(r / 20.0f);
    }
    // denominator = exp(r_iou/T).sum() + exp(r_obj/T) + exp(r_cls/T)
    float eiou = lane < 24 ? e : 0.0f;
    eiou = warp_sum(eiou);
    const float eobj = __shfl_sync(0xffffffffu, e, 24);
    const float ecls = __shfl_sync(0xffffffffu, e, 25);
    const float den = (eiou + eobj) + ecls;
    const float w = (26.0f * e) / den;
    const float wl = w * loss;
    float tot = lane < 24 ? wl : 0.0f;
    tot = warp_sum(tot);
    const float wobj = __shfl_sync(0xffffffffu, wl, 24);
    const float wcls = __shfl_sync(0xffffffffu, wl, 25);
    if (lane < 24) {
        result54[1 + lane] = wl;  // reg_w * loss_iou
        result54[28 + lane] = w;  // reg_w
        weights_n27[lane] = w;
    }
    if (lane == 24) {
        result54[25] = loss;  // loss_obj
        result54[52] = w;
        weights_n27[24] = w;
    }
    if (lane == 25) {
        result54[26] = loss;  // loss_cls
        result54[53] = w;
        weights_n27[25] = w;
    }
    if (lane == 0) {
        result54[0] = ((tot + wobj) + wcls) + 0.0f;
        result54[27] = nfg / ngt;
        weights_n27[26] = nfg;
    }
    if (lane < 26) state26[lane] = loss;
}

// pair value by one warp (lanes over rays, fixed tree); every lane returns it, `l_out` is the lane's ray loss
__device__ __forceinline__ float warp_pair_value(const float* __restrict__ rec, const float* __restrict__ row, float& l_out) {
    const int lane = threadIdx.x & 31;
    const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], row[0], row[1]);
    float l = 0.0f;
    if (lane < P24_RAYS) l = ray_loss(rec[GT_RG + lane], row[2 + lane], d);
    l_out = l;
    return (warp_sum(l) / 24.0f) / 2.0f;
}

// Anchor claimed by several GTs none of which is valid for it (every claim came from a spill): argmin of the
// PENALISED cost over all GTs (losses.py:471-476), first index on ties.  One warp, rare.
__device__ __noinline__ int resolve_conflict(const Params& p, const float* recs, int n, const float* row, int a) {
    const float eo1 = 1.0f + expf(-row[26]);
    const float neg = warp_cls_neg_sum(row + 27, p.nc, eo1);
    const float obj_sig = 1.0f / eo1;
    KV best = {P24_POS_INF, 0x7fffffff};
    for (int g = 0; g < n; ++g) {
        const float* rec = recs + g * GT_REC;
        float l;
        const float v = warp_pair_value(rec, row, l);
        const float c = p24_cost(cls_cost_from(neg, row[27 + gt_class(rec, p.nc)], obj_sig), v, false);
        if (kv_lt(c, g, best.v, best.i)) {
            best.v = c;
            best.i = g;
        }
    }
    (void)a;
    return best.i != 0x7fffffff ? best.i : 0;
}

// Anchor claimed by several GTs: the GT with the smallest cost among those the anchor is valid for (in window and in
// polygon), first index on ties; -1 when there is none.  One warp, lanes over the GTs: the anchor's slot in a GT's
// window table follows from its grid cell and the table's origins.
__device__ __forceinline__ int valid_argmin(const Params& p, int b, int n, int a) {
    const int lane = threadIdx.x & 31;
    int l = 0;
#pragma unroll
    for (int q = 1; q < P24_MAX_LEVELS; ++q) l += (q < p.nlev && a >= p.lev[q].off) ? 1 : 0;
    const int r = a - p.lev[l].off;
    const int iy = r / p.lev[l].W, ix = r - iy * p.lev[l].W;
    KV best = {P24_POS_INF, 0x7fffffff};
    for (int g = lane; g < n; g += 32) {
        const float* tab = p.wtab + ((long long)b * p.Lmax + g) * P24_WT_STRIDE;
        const int sx = ix - __float_as_int(tab[P24_WT_HDR + 2 * l]), sy = iy - __float_as_int(tab[P24_WT_HDR + 2 * l + 1]);
        if (sx >= 0 && sx < P24_WSIDE && sy >= 0 && sy < P24_WSIDE) {
            const float c = tab[l * P24_WSLOTS + sy * P24_WSIDE + sx];
            if (c < P24_POS_INF && kv_lt(c, g, best.v, best.i)) {
                best.v = c;
                best.i = g;
            }
        }
    }
    best = warp_select<false>(best);
    return best.i != 0x7fffffff ? best.i : -1;
}

#define MBOX_SLOT 64  // floats per (epoch half, rank) slot of a mailbox: 28 sums, flag at [32]
#define FIX_SCALE 68719476736.0  // 2^36: fixed-point unit of the loss accumulators (order-independent sums)
#define RESOLVE_GRID_X 32

__device__ __forceinline__ long long to_fix(double x) { return __double2ll_rn(x * FIX_SCALE); }

__global__ void __launch_bounds__(P24_THREADS, 3) k_resolve_loss(const __grid_constant__ Params p) {
    TMARK(2, blockIdx.y * gridDim.x + blockIdx.x, 0);
    pdl_trigger();  // the next step's k_gt_prep may become resident (it waits for this grid's completion)
    pdl_wait();
    TMARK(2, blockIdx.y * gridDim.x + blockIdx.x, 1);
    const int b = blockIdx.y, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int n = p.num_gt[b];
    __shared__ long long s_acc[P24_WARPS][26];
    __shared__ float s_sums[28];
    __shared__ bool s_last;

    const float* img = p.outputs + (long long)b * p.img_stride;
    const float* recs = p.gt_rec + (long long)b * p.Lmax * GT_REC;
    const int nclaim = min(p.nclaimed[b], P24_TOPK * p.Lmax);
    if (blockIdx.x == 0 && tid == 0) p.num_fg[b] = nclaim;  // every claimed anchor ends up foreground (losses.py:479)

    // ---- one warp per claimed anchor: conflict resolution, outputs, loss terms (lanes over rays / classes).
    // Contributions are accumulated as 2^-36 fixed-point integers: the sums do not depend on the list order. ------
    long long acc = 0;  // lane k < 24: sum of loss24[:, k]; lane 24: -sum of obj logits at fg; lane 25: cls BCE
    for (int e = blockIdx.x * P24_WARPS + warp; e < nclaim; e += gridDim.x * P24_WARPS) {
        const int aa = p.claimed[(long long)b * P24_TOPK * p.Lmax + e];
        const long long o = (long long)b * p.A + aa;
        const float* row = img + (long long)aa * p.row_stride;
        const int cnt = p.claim_cnt[o];
        int g = p.claim_gt[o];
        if (cnt > 1) {
            // claimed by several GTs: argmin of the cost over ALL GTs (losses.py:471-476).  Valid pairs always beat
            // penalised ones and their costs are in the GTs' window tables; without any valid pair (every claim came
            // from a spill) the penalised costs are evaluated here
            const int gv = valid_argmin(p, b, n, aa);
            g = gv >= 0 ? gv : resolve_conflict(p, recs, n, row, aa);
        }
        const float* rec = recs + g * GT_REC;
        float l;
        const float v = warp_pair_value(rec, row, l);  // pair value == pred_ious_this_matching (losses.py:491)
        if (lane == 0) {
            p.fg_mask[o] = 1;
            p.matched_gt[o] = g;
            p.pred_iou[o] = v;
        }
        double contrib = (double)l;
        if (p.sums28) {
            // sum_j BCEWithLogits(x_j, t_j), t = v at the GT class and 0 elsewhere (losses.py:246-248, 298-302):
            // sum_j softplus(x_j) - x_c * v, the softplus sum in product form (one log per anchor)
            const int c = gt_class(rec, p.nc);
            float prod = 1.0f, big = 0.0f;
            for (int j = lane; j < p.nc; j += 32) {
                const float x = row[27 + j];
                if (x < 8.0f) prod *= 1.0f + __expf(x);
                else big += x + log1pf(expf(-x));
            }
            prod = warp_prod(prod);
            big = warp_sum(big);
            if (lane == 24) contrib = -(double)row[26];
            if (lane == 25) contrib = ((double)logf(prod) + (double)big) - (double)row[27 + c] * (double)v;
        }
        if (lane < 26) acc += to_fix(contrib);
    }
    TMARK(2, blockIdx.y * gridDim.x + blockIdx.x, 4);
    if (lane < 26) s_acc[warp][lane] = acc;
    __syncthreads();
    if (tid < 26 && p.sums28) {
        long long t = 0;
#pragma unroll
        for (int w = 0; w < P24_WARPS; ++w) t += s_acc[w][tid];
        if (t != 0) atomicAdd((unsigned long long*)&p.acc_fix[b * 28 + tid], (unsigned long long)t);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        // two-level completion count (image, then batch): few atomics per address
        bool last = false;
        const unsigned done = atomicAdd(&p.ticket[1 + b], 1u);
        if (done == (unsigned)gridDim.x - 1u) {
            p.ticket[1 + b] = 0u;  // ready for the next call
            __threadfence();
            const unsigned done2 = atomicAdd(&p.ticket[0], 1u);
            last = (done2 == (unsigned)p.B - 1u);
        }
        s_last = last;
    }
    __syncthreads();
    TMARK(2, blockIdx.y * gridDim.x + blockIdx.x, 5);
    if (!s_last) return;
    __threadfence();
    if (!p.sums28) {  // assignment only (get_assignments): nothing to reduce, but the tickets must be reset
        if (tid == 0) {
            p.ticket[0] = 0u;
            p.ticket[1 + p.B] = 0u;
            p.ticket[2 + p.B] = 0u;
        }
        return;
    }
    // ---- last CTA: batch sums (integer adds over the images: exact), the all-anchor objectness term in a fixed
    // order, then the optional finalize ------------------------------------------------------------------------------
    if (tid < 26) {
        long long t = 0;
        for (int i = 0; i < p.B; ++i) {
            t += __ldcg(&p.acc_fix[i * 28 + tid]);
            p.acc_fix[i * 28 + tid] = 0;  // ready for the next call
        }
        s_sums[tid] = (float)((double)t / FIX_SCALE);
    } else if (tid == 26 || tid == 27) {
        int t = 0;
        const int32_t* src = (tid == 26) ? p.nclaimed : p.num_gt;
        for (int i = 0; i < p.B; ++i) t += min(__ldcg(src + i), tid == 26 ? P24_TOPK * p.Lmax : 0x7fffffff);
        s_sums[tid] = (float)t;
    }
    double objsum = 0.0;
    if (warp == 1) {
        const int nblk = p.B * p.tiles;
        for (int i = lane; i < nblk; i += 32) objsum += __ldcg(p.obj_part + i);
        objsum = warp_sum_d(objsum);
    }
    __syncthreads();
    if (tid == 32) s_sums[24] = (float)((double)s_sums[24] + objsum);
    __syncthreads();
    if (p.nranks > 1) {
        // ---- all-reduce of the 28 sums over peer memory: my sums into everybody's mailbox, a flag behind them, then the
        // contributions of all ranks from my own mailbox, added in rank order (the same bits on every rank).  Two
        // mailbox halves alternate with the epoch: a rank cannot run two epochs ahead of a peer, so a half is never
        // overwritten before everybody has read it.
        const unsigned ep = p.epoch;
        const int half = (int)(ep & 1u) * P24_MAX_RANKS;
        for (int q = warp; q < p.nranks; q += P24_WARPS)
            if (lane < 28) p.mbox[q][(half + p.rank) * MBOX_SLOT + lane] = s_sums[lane];
        __threadfence_system();
        __syncthreads();
        if (tid < p.nranks) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned*>(p.mbox[tid] + (half + p.rank) * MBOX_SLOT + 32) = ep;
            volatile unsigned* f = reinterpret_cast<volatile unsigned*>(p.mbox[p.rank] + (half + tid) * MBOX_SLOT + 32);
            const long long t0 = clock64();
            while (*f != ep) {
                if (clock64() - t0 > 400000000LL) {  // ~0.2 s: a peer is gone; do not hang the GPU
                    atomicOr(p.err_flag, 4);
                    break;
                }
            }
            __threadfence_system();
        }
        __syncthreads();
        if (tid < 28) {
            float t = 0.0f;
            for (int r = 0; r < p.nranks; ++r)
                t += *reinterpret_cast<volatile float*>(p.mbox[p.rank] + (half + r) * MBOX_SLOT + tid);
            s_sums[tid] = t;
        }
        __syncthreads();
    }
    if (tid < 28) p.sums28[tid] = s_sums[tid];
    if (tid == 0) {
        p.ticket[0] = 0u;  // ready for the next call
        p.ticket[1 + p.B] = 0u;
        p.ticket[2 + p.B] = 0u;
    }
    if (p.state26 && warp == 0) finalize_warp(s_sums, p.state26, p.result54, p.weights27);
    TMARK(2, blockIdx.y * gridDim.x + blockIdx.x, 6);
}

__global__ void k_finalize(const float* __restrict__ sums28, float* __restrict__ state26, float* __restrict__ result54,
                           float* __restrict__ weights_n27) {
    finalize_warp(sums28, state26, result54, weights_n27);
}

size_t anchor_pass_smem(int Lmax) { return (size_t)Lmax * GT_REC_HEAD * sizeof(float); }

inline void prof_mark(int i, cudaStream_t st) { p24::prof_mark(i, st); }
#define g_prof_on (p24::prof_on())

template <typename K>
cudaError_t launch(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, const Params& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

}  // namespace

// -------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------
extern "C" size_t p24_workspace_bytes(int B, int A, int Lmax) {
    if (B <= 0 || A <= 0 || Lmax <= 0) return 0;
    return p24_layout(B, A, Lmax).total;
}

extern "C" int p24_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace || ((uintptr_t)workspace & 255) != 0) return P24_E_BADARG;
    return (int)cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int p24_simota_loss_batch(const float* outputs, int64_t img_stride, int64_t row_stride, int B, int A,
                                     int num_classes, const float* labels, int64_t lab_img_stride,
                                     int64_t lab_row_stride, int Lmax, const float* x_shifts, const float* y_shifts,
                                     const float* strides, const int32_t* h_levels, int n_levels, uint8_t* fg_mask,
                                     int32_t* matched_gt, float* pred_iou, int32_t* num_fg, int32_t* num_gt,
                                     int32_t* dyn_k, float* sums28, float* state26, float* result54,
                                     float* weights_n27, void* workspace, size_t workspace_bytes, uint32_t flags,
                                     void* const* h_mailboxes, int rank, int nranks, uint32_t epoch, void* stream) {
    if (!outputs || !labels || !x_shifts || !y_shifts || !strides || !fg_mask || !matched_gt || !pred_iou || !num_fg ||
        !num_gt || !dyn_k || !workspace || !h_levels)
        return P24_E_BADARG;
    if (B <= 0 || A <= 0 || Lmax <= 0 || num_classes <= 0 || Lmax > 65535 || B > 65535) return P24_E_BADARG;
    if (state26 && (!sums28 || !result54 || !weights_n27)) return P24_E_BADARG;
    if (n_levels <= 0) return P24_E_BADARG;
    if (n_levels > P24_MAX_LEVELS) return P24_E_UNSUPPORTED;
    const P24Workspace L = p24_layout(B, A, Lmax);
    if (workspace_bytes < L.total) return P24_E_WORKSPACE;
    if (((uintptr_t)workspace & 255) != 0) return P24_E_BADARG;
    const size_t dyn = anchor_pass_smem(Lmax);
    if (dyn > 160 * 1024 || p24_tiles(A) > MAX_TILES) return P24_E_UNSUPPORTED;
    char* ws = (char*)workspace;
    Params p;
    p.outputs = outputs; p.img_stride = img_stride; p.row_stride = row_stride;
    p.B = B; p.A = A; p.nc = num_classes;
    p.labels = labels; p.lab_img_stride = lab_img_stride; p.lab_row_stride = lab_row_stride; p.Lmax = Lmax;
    p.x_shifts = x_shifts; p.y_shifts = y_shifts; p.strides = strides;
    p.fg_mask = fg_mask; p.matched_gt = matched_gt; p.pred_iou = pred_iou;
    p.num_fg = num_fg; p.num_gt = num_gt; p.dyn_k = dyn_k; p.sums28 = sums28;
    p.state26 = state26; p.result54 = result54; p.weights27 = weights_n27;
    p.gt_rec = (float*)(ws + L.gt_rec);
    p.clist = (float4*)(ws + L.clist);
    p.sval = (float*)(ws + L.sval);
    p.wtab = (float*)(ws + L.wtab);
    p.tbox = (float*)(ws + L.tbox);
    p.seg = (float*)(ws + L.seg);
    p.ccount = (int*)(ws + L.ccount);
    p.claim_cnt = (int*)(ws + L.claim_cnt);
    p.claim_gt = (int*)(ws + L.claim_gt);
    p.obj_part = (double*)(ws + L.obj_part);
    p.claimed = (int*)(ws + L.claimed);
    p.nclaimed = (int*)(ws + L.nclaimed);
    p.acc_fix = (long long*)(ws + L.acc_fix);
    p.ticket = (unsigned*)(ws + L.ticket);
    p.err_flag = (int*)(ws + L.err_flag);
    p.flags = flags;
    p.rank = 0; p.nranks = 1; p.epoch = 0;
    for (int r = 0; r < P24_MAX_RANKS; ++r) p.mbox[r] = nullptr;
    if (h_mailboxes && nranks > 1) {
        if (nranks > P24_MAX_RANKS || rank < 0 || rank >= nranks || !sums28) return P24_E_BADARG;
        for (int r = 0; r < nranks; ++r) {
            if (!h_mailboxes[r]) return P24_E_BADARG;
            p.mbox[r] = (float*)h_mailboxes[r];
        }
        p.rank = rank; p.nranks = nranks; p.epoch = epoch;
    }
    p.tiles = p24_tiles(A);
    p.nlev = n_levels;
    {
        // the levels must tile [0, A) in order: anchors [off, off + W * H) of level l form a W x H grid
        long long next = 0;
        for (int l = 0; l < P24_MAX_LEVELS; ++l) {
            p.lev[l].off = 0; p.lev[l].W = 1; p.lev[l].H = 0; p.lev[l].pad = 0;
            if (l < n_levels) {
                const int32_t* d = h_levels + 4 * l;
                if (d[0] != next || d[1] <= 0 || d[2] <= 0) return P24_E_BADARG;
                p.lev[l].off = d[0]; p.lev[l].W = d[1]; p.lev[l].H = d[2];
                next += (long long)d[1] * d[2];
            }
        }
        if (next != A) return P24_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;

    if (p24::dev_once(1u << 0)) cudaFuncSetAttribute(k_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const bool pdl = !(flags & P24_F_NO_PDL) && !g_prof_on;
    cudaError_t e = cudaSuccess;
    const int n_sm = p24::dev_info().n_sm;
    prof_mark(0, st);
    e = launch(k_gt_prep, dim3(B), dim3(PREP_THREADS), 0, st, pdl, p);
    if (e != cudaSuccess) return (int)e;
    {
        const long long items = (long long)B * p.tiles + (long long)B * Lmax * n_levels;
        const long long cap = 4LL * n_sm;  // one wave of persistent CTAs (5 per SM at 48 registers measured slower)
        e = launch(k_pass, dim3((unsigned)(items < cap ? items : cap)), dim3(P24_THREADS), dyn, st, pdl, p);
        if (e != cudaSuccess) return (int)e;
    }
    prof_mark(1, st);
    e = launch(k_match, dim3(B, Lmax), dim3(MATCH_THREADS), 0, st, pdl, p);
    if (e != cudaSuccess) return (int)e;
    prof_mark(2, st);
    {
        const int gx = (P24_TOPK * Lmax + P24_WARPS - 1) / P24_WARPS;
        e = launch(k_resolve_loss, dim3(gx < RESOLVE_GRID_X ? gx : RESOLVE_GRID_X, B), dim3(P24_THREADS), 0, st, pdl, p);
        if (e != cudaSuccess) return (int)e;
    }
    prof_mark(3, st);
    return (int)cudaGetLastError();
}

extern "C" size_t p24_comm_mailbox_bytes(void) { return (size_t)2 * P24_MAX_RANKS * MBOX_SLOT * sizeof(float); }

extern "C" int p24_comm_alloc(void** d_mailbox) {
    if (!d_mailbox) return P24_E_BADARG;
    cudaError_t e = cudaMalloc(d_mailbox, p24_comm_mailbox_bytes());
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*d_mailbox, 0, p24_comm_mailbox_bytes());
    if (e != cudaSuccess) return (int)e;
    return (int)cudaDeviceSynchronize();
}

extern "C" int p24_comm_free(void* d_mailbox) { return (int)cudaFree(d_mailbox); }

extern "C" int p24_comm_export(void* d_mailbox, void* h_handle64) {
    if (!d_mailbox || !h_handle64) return P24_E_BADARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    return (int)cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(h_handle64), d_mailbox);
}

extern "C" int p24_comm_import(const void* h_handle64, void** d_peer_mailbox) {
    if (!h_handle64 || !d_peer_mailbox) return P24_E_BADARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, sizeof(h));
    return (int)cudaIpcOpenMemHandle(d_peer_mailbox, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" int p24_comm_close(void* d_peer_mailbox) { return (int)cudaIpcCloseMemHandle(d_peer_mailbox); }

extern "C" int p24_loss_finalize(const float* sums28, float* state26, float* result54, float* weights_n27,
                                 void* stream) {
    if (!sums28 || !state26 || !result54 || !weights_n27) return P24_E_BADARG;
    k_finalize<<<1, 32, 0, (cudaStream_t)stream>>>(sums28, state26, result54, weights_n27);
    return (int)cudaGetLastError();
}

#ifdef P24_TIMING
extern "C" int p24_debug_read_timers(unsigned long long* h_out) {
    return (int)cudaMemcpyFromSymbol(h_out, g_tstamp, sizeof(unsigned long long) * 6 * 4096 * 20);
}
#endif
