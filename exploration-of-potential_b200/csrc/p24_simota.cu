// p24_simota.cu — the fused YOLOX-24p SimOTA assignment + loss-sum path for sm_100a.
//
// Replaces, for a whole batch and without ever writing the G x A cost matrix to memory:
//   Loss_Function.get_assignments / get_in_boxes_info / pts_in_poly   models/losses.py:359-592
//   utils.boxes.bboxes_iou + pairwise circle_inter                   utils/boxes.py:102-243
//   Loss_Function.dynamic_k_matching                                 models/losses.py:444-494
//   the loss sums and re-weighting of Loss_Function.forward          models/losses.py:246-345
//
// Kernel chain (caller's stream, no host synchronisation, programmatic dependent launch between the stages):
//   k_prep   one CTA per image: nlabel (losses.py:190) and the per-GT records (ray lengths, inscribed / reject radii)
//   k_pass   persistent CTAs drawing work items from ticket counters:
//            - seed items, one per GT: a handful of anchors that are certainly candidates (centre windows, inscribed
//              discs and polygon tips of the farthest other GTs) are evaluated; the 10th best value T is a certified
//              lower bound of the GT's 10th largest pair value, and far2 the squared centre distance below which no
//              prediction whatsoever can reach T (the bound H* depends on the GT and the distance only);
//            - anchor tiles (256 anchors): THE pass over the head output (cp.async reads of the 27 geometry channels
//              of every row).  Candidate mask (polygon test OR centre window) with geometric pruning and an atan2-free
//              angle test; for the few (GT, candidate) pairs beyond far2 the exact pair value, appended to the GT's
//              top-10 list when it reaches T; sum BCEWithLogits(obj, 0); outputs initialised to background;
//            - centre-window items, one per (GT, level): polygon test, exact pair value and SimOTA cost of the <= 25
//              anchors that can pass the centre-window test -> the GT's window cost table
//   k_tail   one CTA per image: dynamic k = clamp(int(sum of the 10 largest list values), 1) per GT (exact: the list
//            holds every candidate value >= T), the k smallest costs of the window table -> claims in shared memory,
//            conflict resolution (argmin over all GTs), fg_mask / matched_gt / pred_iou, the loss terms of the
//            foreground anchors; the last CTA reduces the batch sums and applies the normalisation and the stateful
//            re-weighting (losses.py:280-345), or hands the sums to the fused peer all-reduce (k_fin)
//   k_fin    (several GPUs) one warp: waits for the peers' sums in the mailbox, adds them in rank order, finalizes.
//            It overlaps the next step's k_prep / k_pass, which do not depend on the global sums.
//
// Compile with -fmad=false: the discrete decisions hang on fp32 thresholds evaluated in the
// reference's operation order; bounds and fast paths use explicit fmaf.
#include "p24_common.cuh"
#include "p24_host.h"
#include <string.h>

namespace {

struct Level {
    int off, W, H;  // anchors [off, off + W * H) form a W x H grid, row-major (yolo_head_24p.py:222-230)
    float st;       // the level's stride (expanded_strides of its anchors)
};

struct Params {
    const float* outputs;            // decoded head output [B, A, 27 + nc] (rows), or NULL:
    long long img_stride, row_stride;
    const float* raw[3][P24_MAX_LEVELS];   // raw conv outputs per level (reg [B,26,H,W], obj [B,1,H,W], cls [B,nc,H,W]) ...
    long long raw_bs[3][P24_MAX_LEVELS];   // ... and their batch strides; decoded on load (yolo_head_24p.py:212-237)
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
    unsigned* ticket;
    long long* acc_fix;
    int* status;
    int* seed_done;
    int* ncand;
    int* lcount;
    float* gt_rec;
    float* wtab;
    float2* list;
    unsigned* cbits;
    int2* wlist;
    float* brute;
    int* claimg;
    int* kreq;
    int* ntake;
    int* rare;
    unsigned flags;
    float* mbox[P24_MAX_RANKS];  // peer mailboxes of the fused all-reduce (nranks > 1)
    int rank, nranks;
    unsigned epoch;
    int tiles;
    int nlev;
    Level lev[P24_MAX_LEVELS];
};

// Debug-only phase timers (-DP24_TIMING): %globaltimer stamps at phase boundaries, read back with p24_debug_read_timers
#ifdef P24_TIMING
#define TM_ROWS 8192
#define TM_SLOTS 16
__device__ unsigned long long g_tstamp[3][TM_ROWS][TM_SLOTS];
__device__ __forceinline__ void tmark(int kern, int row, int slot) {
    if (row >= 0 && row < TM_ROWS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_tstamp[kern][row][slot] = t;
    }
}
#define TMARK(kern, row, slot) do { if ((threadIdx.x & 31) == 0) tmark(kern, row, slot); } while (0)
#define TMARK0(kern, row, slot) do { if (threadIdx.x == 0) tmark(kern, row, slot); } while (0)
#else
#define TMARK(kern, row, slot)
#define TMARK0(kern, row, slot)
#endif

#define TK_ITEM 0   // ticket words
#define TK_LEFF 1
#define TK_TAIL 2
#define TK_SEED 3
#define TK_WIN 4    // window-chunk queue head
#define TK_WTOT 5   // centre-window pairs of the batch (k_prep)

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

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
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

// rank of the lane's value among the 32 lanes' values (0 = largest when MAX; ties -> lower lane first)
template <bool MAX>
__device__ __forceinline__ int lane_rank(float v) {
    const int lane = threadIdx.x & 31;
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float o = __shfl_sync(0xffffffffu, v, j);
        rank += (MAX ? kv_gt(o, j, v, lane) : kv_lt(o, j, v, lane)) ? 1 : 0;
    }
    return rank;
}

// -------------------------------------------------------------------------------------------
// k_prep: nlabel and the per-GT records
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
        rec[GT_FAR2] = 0.0f;   // set by the GT's seed item
        rec[GT_CLS] = row[0];
        rec[GT_RGMAX] = rgmax;
        rec[GT_RGMIN] = rgmin;
        rec[GT_RGMS] = rg2sum * (1.0f / 24.0f);
        rec[GT_RGMEAN] = rgsum * (1.0f / 24.0f);
        rec[GT_T] = P24_NEG_INF;
        rec[59] = 0.0f;
    }
}

// first cell of the 7-wide block that contains every cell centre within 2.5 strides of c (one cell of slack per side)
__device__ __forceinline__ int window_origin(float c, float st) {
    float v = floorf(c / st) - 3.0f;
    v = fminf(fmaxf(v, -1.0e6f), 1.0e6f);  // NaN -> -1e6: an empty window
    return (int)v;
}

// which of the lane's slots s = lane + 32 q of a GT's window table pass the strict centre-window test (losses.py:523-542,
// in the reference's arithmetic); bit q of the result.  The anchor of a slot follows from the level grids.
__device__ __forceinline__ unsigned window_slot_mask(const Params& p, float gcx, float gcy) {
    const int lane = threadIdx.x & 31;
    unsigned m = 0u;
#pragma unroll
    for (int q = 0; q < (P24_WT_HDR + 31) / 32; ++q) {
        const int s = lane + 32 * q;
        const int l = s / P24_WSLOTS, r = s - l * P24_WSLOTS;
        if (l < p.nlev) {
            const float st = p.lev[l].st;
            const int sy = r / P24_WSIDE, sx = r - sy * P24_WSIDE;
            const int ix = window_origin(gcx, st) + sx, iy = window_origin(gcy, st) + sy;
            // (the grid is the head's, validated by the host side: x_shift = column, y_shift = row, one stride per level)
            if (ix >= 0 && ix < p.lev[l].W && iy >= 0 && iy < p.lev[l].H &&
                p24_in_centre(gcx, gcy, p24_anchor_centre((float)ix, st), p24_anchor_centre((float)iy, st), st))
                m |= 1u << q;
        }
    }
    return m;
}

// -------------------------------------------------------------------------------------------
// shared device helpers
// -------------------------------------------------------------------------------------------
// Out-of-line copies of the two heavy scalar routines: the kernels call them from many places
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

__device__ __forceinline__ int gt_class(const float* rec, int nc) {
    const int c = (int)rec[GT_CLS];
    return min(max(c, 0), nc - 1);
}

// Sum over all classes of BCE(p_j, 0) (losses.py:406-416)
// ... by the 8 lanes of a group, term by term (fallback of the product form)
__device__ __noinline__ float group_cls_neg_sum(const float* __restrict__ cls, int nc, float eo1, unsigned m) {
    const int sub = threadIdx.x & 7;
    const float obj_sig = 1.0f / eo1;
    float s = 0.0f;
    for (int j = sub; j < nc; j += 8) s += p24_bce_neg(p24_joint_prob(cls[j], obj_sig));
    return group_sum(s, m);
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
// Access to one anchor's channels of the head output.  Two layouts:
//   rows    the decoded buffer [B, A, 27 + nc] that YOLOXHead.forward(train=True) returns (yolo_head_24p.py:197);
//   raw     the head's raw per-level conv outputs reg [B,26,H,W], obj [B,1,H,W], cls [B,nc,H,W] (yolo_head_24p.py:160-164),
//           decoded on load exactly as get_output_and_grid does (yolo_head_24p.py:233-235): centre (v + grid) * stride,
//           radii exp(v) * stride; the cat / view / permute / reshape copies of the reference never happen.
// -------------------------------------------------------------------------------------------
struct Src {
    const float* reg;   // channel c of the anchor at reg[c * plane]
    const float* obj;
    const float* cls;   // class j at cls[j * plane]
    long long plane;    // 1 (rows) or W * H of the level (raw)
    float gx, gy, st;   // grid cell and stride (raw)
    bool raw;
};

__device__ __forceinline__ Src src_of(const Params& p, int b, int a) {
    Src s;
    if (p.outputs) {
        const float* row = p.outputs + (long long)b * p.img_stride + (long long)a * p.row_stride;
        s.reg = row;
        s.obj = row + 26;
        s.cls = row + 27;
        s.plane = 1;
        s.gx = s.gy = 0.0f;
        s.st = 1.0f;
        s.raw = false;
        return s;
    }
    int l = 0;
#pragma unroll
    for (int q = 1; q < P24_MAX_LEVELS; ++q) l += (q < p.nlev && a >= p.lev[q].off) ? 1 : 0;
    const int r = a - p.lev[l].off;
    const int iy = r / p.lev[l].W, ix = r - iy * p.lev[l].W;
    s.plane = (long long)p.lev[l].W * p.lev[l].H;
    s.reg = p.raw[0][l] + (long long)b * p.raw_bs[0][l] + r;
    s.obj = p.raw[1][l] + (long long)b * p.raw_bs[1][l] + r;
    s.cls = p.raw[2][l] + (long long)b * p.raw_bs[2][l] + r;
    s.gx = (float)ix;
    s.gy = (float)iy;
    s.st = p.lev[l].st;
    s.raw = true;
    return s;
}
// decoded geometry channel c in [0, 26): centre x, centre y, 24 radii
__device__ __forceinline__ float src_geo(const Src& s, int c) {
    const float v = s.reg[(long long)c * s.plane];
    if (!s.raw) return v;
    if (c >= 2) return expf(v) * s.st;             // output[..., 2:26] = exp(output[..., 2:26]) * stride
    return (v + (c == 0 ? s.gx : s.gy)) * s.st;    // output[..., :2] = (output[..., :2] + grid) * stride
}
__device__ __forceinline__ float src_obj(const Src& s) { return s.obj[0]; }
__device__ __forceinline__ float src_cls(const Src& s, int j) { return s.cls[(long long)j * s.plane]; }

// exact pair value of (GT record, anchor of the head output): utils/boxes.py:166-243, one thread
__device__ __noinline__ float pair_value_src(const float* __restrict__ rec, const Src& s) {
    const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], src_geo(s, 0), src_geo(s, 1));
    float sm = 0.0f;
#pragma unroll 1
    for (int k = 0; k < P24_RAYS; ++k) sm = sm + ray_loss(rec[GT_RG + k], src_geo(s, 2 + k), d);
    return (sm / 24.0f) / 2.0f;
}
// sum over all classes of BCE(p_j, 0) (losses.py:406-416), term by term, by a single thread / an 8-lane group (rare paths)
__device__ __noinline__ float thread_cls_neg_sum_src(const Src& s, int nc, float eo1) {
    const float obj_sig = 1.0f / eo1;
    float t = 0.0f;
    for (int j = 0; j < nc; ++j) t += p24_bce_neg(p24_joint_prob(src_cls(s, j), obj_sig));
    return t;
}
__device__ __noinline__ float group_cls_neg_sum_src(const Src& s, int nc, float eo1, unsigned m) {
    const int sub = threadIdx.x & 7;
    const float obj_sig = 1.0f / eo1;
    float t = 0.0f;
    for (int j = sub; j < nc; j += 8) t += p24_bce_neg(p24_joint_prob(src_cls(s, j), obj_sig));
    return group_sum(t, m);
}

// -------------------------------------------------------------------------------------------
// Upper bound of the pair value as a function of the centre distance d alone.  Any ray has
// loss <= max(1, 2 - 4 (rg^2 + rp^2) / (rg + rp + d)^2) (nested rays: loss <= 1; partial and apart rays:
// loss <= 2 - uni/cs with the apart formula; tests/test_bounds_cpu.py).  Over all rp > 0 the fraction is smallest at
// rp* = rg^2 / (rg + d), where it equals rg^2 / ((rg + d)^2 + rg^2).  So every ray has
// loss <= max(1, 2 - 4 rg^2 / ((rg + d)^2 + rg^2)) whatever the prediction: monotone in d.
// One warp, lanes over the rays; every lane returns the bound H*(d) of the pair value.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_bound_Hstar(float rg_lane, float d) {
    const int lane = threadIdx.x & 31;
    float t = 0.0f;
    if (lane < P24_RAYS) {
        const float q = rg_lane + d;
        t = fmaxf(1.0f, 2.0f - __fdividef(4.0f * rg_lane * rg_lane, fmaf(q, q, rg_lane * rg_lane)));
    }
    return warp_sum(t) * (1.0f / 48.0f);
}

// -------------------------------------------------------------------------------------------
// k_pass, seed items.  Dynamic k needs the 10 LARGEST pair values of a GT over the image's candidate anchors
// (losses.py:452-456); the pair value grows with the centre distance, so they belong to candidates far away from the
// GT.  A seed item evaluates a handful of anchors that are certainly candidates and far away: the centre-window and
// inscribed-disc anchors of the SEED_FAR farthest GTs of the image on the side that looks away from this GT, and the
// anchors just inside their polygon tips on that side (verified with the polygon test).  T = the 10th best seed value
// is a certified lower bound of the 10th largest pair value, and far2 = the squared centre distance below which
// H*(d) < T: the anchor tiles evaluate only the (GT, candidate) pairs beyond it.
// -------------------------------------------------------------------------------------------
#define SEED_FAR 3      // farthest GTs whose centre-window / inscribed-disc anchors serve as seeds

__device__ __forceinline__ int cell_index(float q, float st) {
    float v = floorf(q / st);
    v = fminf(fmaxf(v, -1.0e6f), 1.0e6f);  // NaN -> -1e6: outside every grid
    return (int)v;
}

// One seed point of GT `rec` by one lane: the grid cell of (qx, qy) on level l.  When the cell's anchor is certainly a
// candidate (the very tests of the anchor tiles: inscribed disc / centre window / polygon of GT h) and every ray of the
// pair is in the "apart" branch (the reference's own fp32 comparison), the pair value has the closed form
// (1/48) sum_k (2 - 4 (rg^2 + rp^2) / (rg + rp + d)^2), evaluated in fast arithmetic to within 3e-6: returns
// (value - 1e-5, anchor), a certified LOWER bound of a candidate's pair value; (-inf, -1) otherwise.
__device__ __forceinline__ KV seed_point(const Params& p, const float* __restrict__ rec, const float* __restrict__ h, int b, int l,
                                         float qx, float qy, bool valid) {
    KV out = {P24_NEG_INF, -1};
    const Level lv = p.lev[l];
    const float st = lv.st;
    const int ix = cell_index(qx, st), iy = cell_index(qy, st);
    if (!(valid && ix >= 0 && ix < lv.W && iy >= 0 && iy < lv.H)) return out;
    const int a = lv.off + iy * lv.W + ix;
    // the anchor's row is requested right away: it is in flight while the candidate tests run
    const Src src = src_of(p, b, a);
    float rpv[P24_RAYS];
#pragma unroll
    for (int k = 0; k < P24_RAYS; ++k) rpv[k] = src_geo(src, 2 + k);
    const float pcx = src_geo(src, 0), pcy = src_geo(src, 1);
    // (the grid is the head's, validated by the host side: x_shift = column, y_shift = row, one stride per level)
    const float xc = p24_anchor_centre((float)ix, st), yc = p24_anchor_centre((float)iy, st);
    const float hcx = h[GT_CX], hcy = h[GT_CY];
    const float dx = hcx - xc, dy = hcy - yc;
    bool cand = fmaf(dx, dx, dy * dy) < h[GT_RIN2];
    if (!cand) cand = p24_in_centre(hcx, hcy, xc, yc, st);
    if (!cand) cand = p24_in_polygon(h + GT_VX, h + GT_VY, xc, yc);
    if (!cand) return out;
    const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], pcx, pcy);
    float sm = 0.0f;
    bool apart = true;
#pragma unroll
    for (int k = 0; k < P24_RAYS; ++k) {
        const float rg = rec[GT_RG + k], rp = rpv[k];
        apart = apart && (d >= rg + rp) && (rp >= 0.25f);
        const float t = (rg + rp) + d;
        sm += 2.0f - __fdividef(4.0f * fmaf(rg, rg, rp * rp), t * t);
    }
    // (a seed with a ray that is not apart would need the exact evaluation: skipped, far seeds are apart)
    const float v = sm * (1.0f / 48.0f) - 1e-5f;
    if (apart && v == v) {
        out.v = v;
        out.i = a;
    }
    return out;
}

// -------------------------------------------------------------------------------------------
// k_prep: grid (ceil(Lmax / 2), B); CTA (c, b) prepares the GTs 2c and 2c + 1 of image b, four warps each.
//   part 1 (warp 0 of the GT): nlabel is counted by every CTA (the label block stays in L2), the GT's record, its window
//          table and its centre-window pairs (appended to the batch's list: the window chunks of k_pass);
//   barrier over the image's CTAs (a counter of finished records; see p24_simota_loss_batch for the launch condition);
//   part 2 (warps 0..2 of the GT): the seeds, 32 points per warp: warp 0 / warp 1 the polygon vertices of the image that
//          are far from the GT (every lane the farthest of its share) on the two levels that suit the GT, warp 2 the
//          centre-window / inscribed-disc anchors of the SEED_FAR farthest GTs.  The 10th largest of their certified
//          values is T, and far2 follows from it.
// mode 0: both parts; 1: part 1 only; 2: part 2 only (two launches, when the grid would not be resident at once).
// -------------------------------------------------------------------------------------------
#define PREP_THREADS 256
#define PREP_GTS 2
struct PrepShared {
    int n, base;
    int wcnt[PREP_GTS];
    int far[PREP_GTS][4];
    float tval[PREP_GTS][3][P24_TOPK];
    int tanc[PREP_GTS][3][P24_TOPK];
};

__global__ void __launch_bounds__(PREP_THREADS) k_prep(const __grid_constant__ Params p, int mode) {
    extern __shared__ float4 pr_dyn4[];  // [Lmax * GT_REC] floats: the image's records (part 2)
    __shared__ PrepShared S;
    // launched as a programmatic dependent of whatever precedes it in the stream (in back-to-back steps: the previous
    // step's k_tail, which triggers at once): resident early, it starts the moment that work is complete
    pdl_wait();
    pdl_trigger();
    float* recs = reinterpret_cast<float*>(pr_dyn4);
    const int b = blockIdx.y, c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    TMARK0(0, b * 64 + c, 0);
    const int gi = warp >> 2, role = warp & 3;
    const int g = c * PREP_GTS + gi;
    const float* lab = p.labels + (long long)b * p.lab_img_stride;
    int n;
    if (mode != 2) {
        n = block_count_labels(p, lab, &S.n);
        if (tid == 0 && c == 0) {
            p.num_gt[b] = n;
            p.num_fg[b] = 0;  // k_tail adds the foreground anchors of the image's cluster CTAs
            atomicMax(&p.ticket[TK_LEFF], (unsigned)n);
        }
    } else {
        n = p.num_gt[b];
    }
    if (c * PREP_GTS >= n) return;  // (the whole CTA)
    TMARK0(0, b * 64 + c, 1);
    const bool has = g < n;
    if (mode != 2) {
        // ---- part 1 ------------------------------------------------------------------------------------------------------
        unsigned wmask = 0u;
        float gcx = 0.0f, gcy = 0.0f;
        if (has && role == 0) {
            const float* row = lab + (long long)g * p.lab_row_stride;
            float* rec = p.gt_rec + ((long long)b * p.Lmax + g) * GT_REC;
            warp_gt_record(row, rec);
            gcx = row[1];
            gcy = row[2];
            // the GT's window cost table: origins of its 7 x 7 block of cells per level, every slot "not valid" until the
            // pair's cost is stored
            float* tab = p.wtab + ((long long)b * p.Lmax + g) * P24_WT_STRIDE;
            if (lane < 2 * P24_MAX_LEVELS) {
                const int l = lane >> 1;
                const float o = __int_as_float(l < p.nlev ? window_origin(lane & 1 ? gcy : gcx, p.lev[l].st) : 0);
                rec[GT_ORG + lane] = o;
                tab[P24_WT_HDR + lane] = o;
            }
            for (int s = lane; s < P24_WT_HDR; s += 32) tab[s] = P24_POS_INF;
            wmask = window_slot_mask(p, gcx, gcy);
        }
        const int cnt = warp_sum_i(__popc(wmask));
        if (lane == 0 && role == 0) S.wcnt[gi] = has ? cnt : 0;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < PREP_GTS; ++w) {
                const int cc = S.wcnt[w];
                S.wcnt[w] = tot;
                tot += cc;
            }
            S.base = tot ? (int)atomicAdd(&p.ticket[TK_WTOT], (unsigned)tot) : 0;
        }
        __syncthreads();
        if (has && role == 0) {
            int at = S.base + S.wcnt[gi];
#pragma unroll
            for (int q = 0; q < (P24_WT_HDR + 31) / 32; ++q) {
                const bool in = (wmask >> q) & 1u;
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const int s = lane + 32 * q;
                    const int l = s / P24_WSLOTS, r = s - l * P24_WSLOTS;
                    const int sy = r / P24_WSIDE, sx = r - sy * P24_WSIDE;
                    const int ix = window_origin(gcx, p.lev[l].st) + sx, iy = window_origin(gcy, p.lev[l].st) + sy;
                    p.wlist[at + __popc(bal & ((1u << lane) - 1u))] = make_int2(b * p.Lmax + g, p.lev[l].off + iy * p.lev[l].W + ix);
                }
                at += __popc(bal);
            }
        }
        if (mode == 1) return;
        TMARK0(0, b * 64 + c, 2);
        // ---- the image's records are complete when every CTA of the image has passed here --------------------------------
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            atomicAdd(&p.seed_done[b], min(PREP_GTS, n - c * PREP_GTS));
            while (ld_acquire(&p.seed_done[b]) < n) __nanosleep(32);
        }
        __syncthreads();
        TMARK0(0, b * 64 + c, 3);
    }
    // ---- part 2 ----------------------------------------------------------------------------------------------------------
    {
        const float4* gsrc = reinterpret_cast<const float4*>(p.gt_rec + (long long)b * p.Lmax * GT_REC);
        for (int i = tid; i < n * (GT_REC / 4); i += PREP_THREADS) pr_dyn4[i] = __ldcg(gsrc + i);
    }
    __syncthreads();
    TMARK0(0, b * 64 + c, 4);
    const float* rec = recs + min(g, n - 1) * GT_REC;
    const float gcx = rec[GT_CX], gcy = rec[GT_CY];
    const bool filter = has && !(p.flags & P24_F_NO_FILTER) && rec[GT_RGMIN] >= 0.25f && rec[GT_RGMAX] < 1.0e6f;
    if (filter && role < 3) {
        KV sd = {P24_NEG_INF, -1};
        if (role < 2) {
            // the farthest polygon vertex of the lane's share of the image's vertices
            float best = P24_NEG_INF;
            int bi = -1;
            for (int i = lane; i < n * P24_RAYS; i += 32) {
                const int h = i / P24_RAYS, k = i - h * P24_RAYS;
                const float* r = recs + h * GT_REC;
                const float dx = r[GT_VX + k] - gcx, dy = r[GT_VY + k] - gcy;
                const float d2 = fmaf(dx, dx, dy * dy);
                if (d2 > best) {
                    best = d2;
                    bi = (h << 5) | k;
                }
            }
            const float dfar2 = warp_max(best);
            // the two levels whose typical predicted radius (about 1.1 strides) is closest to the radius that maximises the
            // bound at the far distance, rp* = rg^2 / (rg + d): warp 0 takes the best, warp 1 the second best
            int lvA = 0, lvB = 0;
            {
                const float rstar = __fdividef(rec[GT_RGMS], rec[GT_RGMEAN] + sqrtf(fmaxf(dfar2, 0.0f)) + 1e-6f);
                float eA = P24_POS_INF, eB = P24_POS_INF;
                for (int l = 0; l < p.nlev; ++l) {
                    const float e = fabsf(1.13f * p.lev[l].st - rstar);
                    if (e < eA) {
                        eB = eA;
                        lvB = lvA;
                        eA = e;
                        lvA = l;
                    } else if (e < eB) {
                        eB = e;
                        lvB = l;
                    }
                }
            }
            const int l = role == 0 ? lvA : lvB;
            const float* h = recs + (max(bi, 0) >> 5) * GT_REC;
            const int k = max(bi, 0) & 31;
            const float st = p.lev[l].st;
            // one stride inside the vertex, on the ray from the vertex's own GT centre
            const float rr = h[GT_RG + k];
            const float fct = fmaxf(0.0f, __fdividef(rr - st, fmaxf(rr, 1e-6f)));
            sd = seed_point(p, rec, h, b, l, fmaf(h[GT_VX + k] - h[GT_CX], fct, h[GT_CX]),
                            fmaf(h[GT_VY + k] - h[GT_CY], fct, h[GT_CY]), bi >= 0 && (role == 0 || p.nlev > 1));
        } else {
            // the SEED_FAR farthest GTs (centre distance + their largest ray), this GT included: centre, far end of the
            // inscribed disc, 2 strides out, on every level (up to 3): 27 lanes
            float key = P24_NEG_INF;
            int hb = -1;
            for (int h = lane; h < n; h += 32) {
                const float* r = recs + h * GT_REC;
                const float dx = r[GT_CX] - gcx, dy = r[GT_CY] - gcy;
                const float kk = sqrtf(fmaf(dx, dx, dy * dy)) + r[GT_RGMAX];
                if (kk > key) {
                    key = kk;
                    hb = h;
                }
            }
            const int rk = lane_rank<true>(hb >= 0 ? key : P24_NEG_INF);
            if (rk < SEED_FAR) S.far[gi][rk] = hb;   // (-1: fewer GTs than SEED_FAR)
            __syncwarp();
            const int f = lane / 9, r0 = lane - f * 9;
            const int l = r0 / 3, pt = r0 - l * 3;
            const int hs = f < SEED_FAR ? S.far[gi][f] : -1;
            const float* h = recs + max(hs, 0) * GT_REC;
            const int ll = min(l, p.nlev - 1);
            const float st = p.lev[ll].st;
            float ux = h[GT_CX] - gcx, uy = h[GT_CY] - gcy;
            const float nn = fmaf(ux, ux, uy * uy);
            if (nn > 1e-12f) {
                const float inv = rsqrtf(nn);
                ux *= inv;
                uy *= inv;
            } else {
                ux = 1.0f;
                uy = 0.0f;
            }
            const float rho = pt == 0 ? 0.0f : (pt == 1 ? fmaxf(0.0f, sqrtf(h[GT_RIN2]) - 0.75f * st) : 2.0f * st);
            sd = seed_point(p, rec, h, b, ll, fmaf(ux, rho, h[GT_CX]), fmaf(uy, rho, h[GT_CY]), hs >= 0 && lane < 27 && l < p.nlev);
        }
        if (warp == 0) TMARK(0, b * 64 + c, 5);
        // the warp's 10 largest values over distinct anchors (a copy of an anchor counts once: the lower lane keeps it)
        bool dup = false;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const int aj = __shfl_sync(0xffffffffu, sd.i, j);
            dup = dup || (sd.i >= 0 && aj == sd.i && j < lane);
        }
        const float v = dup ? P24_NEG_INF : sd.v;
        const int rk = lane_rank<true>(v);
        if (rk < P24_TOPK) {
            S.tval[gi][role][rk] = v;
            S.tanc[gi][role][rk] = v > P24_NEG_INF ? sd.i : -1;
        }
    }
    __syncthreads();
    TMARK0(0, b * 64 + c, 6);
    if (has && role == 0) {
        float T = P24_NEG_INF;
        if (filter) {
            // the 10th largest of the three warps' best values over distinct anchors (30 values, one per lane)
            const float* tv = &S.tval[gi][0][0];
            const int* ta = &S.tanc[gi][0][0];
            float v = lane < 3 * P24_TOPK ? tv[lane] : P24_NEG_INF;
            const int a = lane < 3 * P24_TOPK ? ta[lane] : -1;
            bool dup = false;
#pragma unroll 8
            for (int j = 0; j < 32; ++j) {
                const int aj = __shfl_sync(0xffffffffu, a, j);
                dup = dup || (a >= 0 && aj == a && j < lane);
            }
            if (dup || a < 0) v = P24_NEG_INF;
            const int rk = lane_rank<true>(v);
            T = warp_max(rk == P24_TOPK - 1 ? v : P24_NEG_INF);
        }
        // ---- far2: pairs closer than D cannot reach T.  H*(d) + 3e-5 < T is monotone in d: rounds of a 32-way search, every
        // lane evaluating the bound at its own distance ----------------------------------------------------------------------
        float far2 = 0.0f;  // 0: every pair is evaluated
        if (T > P24_NEG_INF) {
            float lo = 0.0f, step = 256.0f;  // pairs farther apart than 32 * 256 px are always evaluated
#pragma unroll 1
            for (int round = 0; round < 3; ++round) {
                const float dd = lo + step * (float)lane;
                float sum = 0.0f;
#pragma unroll
                for (int k = 0; k < P24_RAYS; ++k) {
                    const float rg = rec[GT_RG + k];
                    const float q = rg + dd;
                    sum += fmaxf(1.0f, 2.0f - __fdividef(4.0f * rg * rg, fmaf(q, q, rg * rg)));
                }
                const bool below = sum * (1.0f / 48.0f) + 3e-5f < T;      // true for a prefix of the lanes (monotone)
                const unsigned bal = __ballot_sync(0xffffffffu, below);
                const int nb = bal == 0xffffffffu ? 32 : __ffs(~bal) - 1;  // length of the leading run of lanes
                if (nb == 0) break;  // even lo is not below: D = lo
                lo = lo + step * (float)(nb - 1);
                step = step * (1.0f / 32.0f);
            }
            const float D = fmaxf(0.0f, lo * (1.0f - 1e-4f) - 0.02f);
            far2 = D * D;
            if (!(far2 == far2)) far2 = 0.0f;
        }
        if (lane == 0) {
            float* myrec = p.gt_rec + ((long long)b * p.Lmax + g) * GT_REC;
            myrec[GT_FAR2] = far2;
            myrec[GT_T] = T;
        }
        if (warp == 0) TMARK(0, b * 64 + c, 7);
    }
}

// -------------------------------------------------------------------------------------------
// k_pass, anchor tiles: one item per 256-anchor tile
// -------------------------------------------------------------------------------------------
#define ITEM_CAP 1024
#define ROW_CH 27  // channels 0..26 of a head row are read here: centre, 24 radii, objectness
#define FIX_SCALE 68719476736.0     // 2^36: fixed-point unit of the loss accumulators (order-independent sums)
#define FIX_SCALE_OBJ 4294967296.0  // 2^32 for the objectness sum over all anchors (range for B * A terms)

__device__ __forceinline__ long long to_fix(double x) { return __double2ll_rn(x * FIX_SCALE); }

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
    unsigned items[ITEM_CAP];
    float box[P24_WARPS][9];   // per warp: boxes of the anchor centres and of the predicted centres, largest stride
    int nitems, nnear, nfarl;
};

// the tile's rows into shared memory.  Row layout: each warp reads its 32 rows, 27 contiguous floats per row, with cp.async.
// Raw layout: every lane reads the 27 channels of its own anchor -- 32 consecutive anchors of a level are 32 consecutive
// floats of every channel plane: fully coalesced -- and decodes them on the way.
__device__ __forceinline__ void stage_rows(const Params& p, AnchorShared& S, int b, int tile) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a0 = tile * P24_THREADS + warp * 32;
    if (p.outputs) {
        const float* img = p.outputs + (long long)b * p.img_stride;
        const int nrow = min(32, p.A - a0);
        if (lane < ROW_CH) {
            const float* src = img + (long long)a0 * p.row_stride + lane;
            for (int r = 0; r < nrow; ++r) cp_async4(&S.row[warp][lane][r], src + (long long)r * p.row_stride);
        }
    } else if (a0 + lane < p.A) {
        const Src s = src_of(p, b, a0 + lane);
        float v[ROW_CH];
#pragma unroll
        for (int c = 0; c < 26; ++c) v[c] = s.reg[(long long)c * s.plane];   // all in flight together
        v[26] = src_obj(s);
        S.row[warp][0][lane] = (v[0] + s.gx) * s.st;
        S.row[warp][1][lane] = (v[1] + s.gy) * s.st;
#pragma unroll
        for (int c = 2; c < 26; ++c) S.row[warp][c][lane] = expf(v[c]) * s.st;
        S.row[warp][26][lane] = v[26];
    }
}

// Pair (GT g, staged anchor al) lies beyond far2.  A cheap upper bound of its value (the "apart" closed form bounds every
// ray, p24_ray_loss_ub; for a pair whose rays are all apart -- the reference's own fp32 comparison -- it IS the value to
// within 3e-6); when it reaches T the pair goes to the GT's list as (bound, anchor | all-apart flag): k_tail refines the
// threshold with the certified lower bounds and evaluates the few pairs that remain exactly.
__device__ __noinline__ void far_pair(const Params& p, const float* __restrict__ s_rec, const AnchorShared& S, int b, int tile,
                                      int g, int al) {
    const float* rec = s_rec + g * GT_REC;
    const int wr = al >> 5, lr = al & 31;
    const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], S.row[wr][0][lr], S.row[wr][1][lr]);
    float sm = 0.0f;
    bool apart = true;
#pragma unroll 4
    for (int k = 0; k < P24_RAYS; ++k) {
        const float rg = rec[GT_RG + k], rp = S.row[wr][2 + k][lr];
        apart = apart && (d >= rg + rp);
        sm += p24_ray_loss_ub(rg, rp, d);
    }
    float ub = sm * (1.0f / 48.0f) + 2e-5f;
    if ((S.cand[al] & 2) || !(ub == ub)) {  // tiny predicted radius / NaN: no bound, evaluated exactly by k_tail
        ub = P24_POS_INF;
        apart = false;
    }
    if (ub >= rec[GT_T]) {
        const int slot = b * p.Lmax + g;
        const int at = atomicAdd(&p.lcount[slot], 1);
        if (at < P24_LISTCAP)
            p.list[(long long)slot * P24_LISTCAP + at] =
                make_float2(ub, __int_as_float((tile * P24_THREADS + al) | (apart ? (int)0x80000000 : 0)));
    }
}

// One (GT, anchor) pair that passes the centre-window test of losses.py:523-542, by an 8-lane group: polygon test
// (inscribed-disc accept, else the reference-order edge terms, 3 per lane), exact pair value and SimOTA cost when inside
// -> the GT's window cost table, by slot (level, row, column) of its 7 x 7 block of cells.  The anchor of a slot and the
// slot of an anchor are both computable, which is what k_tail (selection and conflict argmin) relies on.
__device__ __forceinline__ void window_pair(const Params& p, int slot, int aa, unsigned gm) {
    const int sub = threadIdx.x & 7;
    const int b = slot / p.Lmax;
    const float* rec = p.gt_rec + (long long)slot * GT_REC;
    const Src src = src_of(p, b, aa);
    // everything the pair needs is requested up front
    const float gcx = rec[GT_CX], gcy = rec[GT_CY], rin2 = rec[GT_RIN2];
    const int c = gt_class(rec, p.nc);
    float cl[10], rp[3], rg[3];
#pragma unroll
    for (int q = 0; q < 10; ++q) cl[q] = (sub + 8 * q < p.nc) ? src_cls(src, sub + 8 * q) : 0.0f;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        rp[q] = src_geo(src, 2 + sub * 3 + q);
        rg[q] = rec[GT_RG + sub * 3 + q];
    }
    const float pcx = src_geo(src, 0), pcy = src_geo(src, 1), obj = src_obj(src), clsc = src_cls(src, c);
    int l = 0;
#pragma unroll
    for (int q = 1; q < P24_MAX_LEVELS; ++q) l += (q < p.nlev && aa >= p.lev[q].off) ? 1 : 0;
    const int r = aa - p.lev[l].off;
    const int iy = r / p.lev[l].W, ix = r - iy * p.lev[l].W;
    const float st = p.lev[l].st;
    // (the grid is the head's, validated by the host side: x_shift = column, y_shift = row, one stride per level)
    const float xc = p24_anchor_centre((float)ix, st), yc = p24_anchor_centre((float)iy, st);
    bool inside = true;
    {
        // inside the inscribed disc the angle sum is >= 360 (see warp_gt_record): no edge terms needed
        const float ddx = gcx - xc, ddy = gcy - yc;
        if (!(fmaf(ddx, ddx, ddy * ddy) < rin2) || (p.flags & P24_F_NO_PRUNE)) {
            float ang = 0.0f;
#pragma unroll 1
            for (int q = 0; q < 3; ++q) {
                const int k = sub * 3 + q;
                const int k2 = (k == P24_RAYS - 1) ? 0 : k + 1;
                ang = ang + edge_angle(rec[GT_VX + k] - xc, rec[GT_VY + k] - yc, rec[GT_VX + k2] - xc, rec[GT_VY + k2] - yc);
            }
            ang = group_sum(ang, gm);
            inside = ang >= 350.0f;  // losses.py:588
        }
    }
    if (!inside) return;
    const float d = p24_centre_dist(gcx, gcy, pcx, pcy);
    float sm = 0.0f;
#pragma unroll
    for (int q = 0; q < 3; ++q) sm = sm + ray_loss(rg[q], rp[q], d);
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
        neg = (prod > 1e-30f) ? (-logf(prod) + 100.0f * (float)nsat) : group_cls_neg_sum_src(src, p.nc, eo1, gm);
    } else {
        // many classes: the log of a per-lane product, restarted before it can underflow
        float prod = 1.0f, lsum = 0.0f;
        int nsat = 0;
        for (int j = sub; j < p.nc; j += 8) {
            p24_neg_factor(src_cls(src, j), eo1, prod, nsat);
            if (prod < 1e-20f) {
                lsum += logf(prod);
                prod = 1.0f;
            }
        }
        lsum += logf(prod);
        lsum = group_sum(lsum, gm);
        nsat = group_sum_i(nsat, gm);
        neg = -lsum + 100.0f * (float)nsat;
        if (!(neg == neg) || neg == P24_POS_INF) neg = group_cls_neg_sum_src(src, p.nc, eo1, gm);
    }
    float cost = p24_cost(cls_cost_from(neg, clsc, 1.0f / eo1), v, true);
    if (!(cost < 3.0e38f)) cost = 3.0e38f;  // NaN / inf inputs: keep the pair selectable, last
    const int sx = ix - __float_as_int(rec[GT_ORG + 2 * l]), sy = iy - __float_as_int(rec[GT_ORG + 2 * l + 1]);
    if (sub == 0 && sx >= 0 && sx < P24_WSIDE && sy >= 0 && sy < P24_WSIDE)
        p.wtab[(long long)slot * P24_WT_STRIDE + l * P24_WSLOTS + sy * P24_WSIDE + sx] = cost;
}

__device__ __forceinline__ void anchor_part(const Params& p, float* s_rec, AnchorShared& S, int b, int tile, bool staged) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int a = tile * P24_THREADS + tid;
    const bool active = a < p.A;
    if (!staged) stage_rows(p, S, b, tile);
    float st = 1.f, xs = 0.f, ys = 0.f;
    if (active) {
        st = p.strides[a];
        xs = p.x_shifts[a];
        ys = p.y_shifts[a];
    }
    if (tid == 0) S.nitems = 0;
    S.cand[tid] = 0;
    const int n = p.num_gt[b];
    TMARK0(1, b * p.tiles + tile, 1);
    __syncthreads();
    {
        const float4* gsrc = reinterpret_cast<const float4*>(p.gt_rec + (long long)b * p.Lmax * GT_REC);
        float4* dst = reinterpret_cast<float4*>(s_rec);
        for (int i = tid; i < n * (GT_REC / 4); i += P24_THREADS) dst[i] = __ldcg(gsrc + i);
    }
    cp_async_wait_all();
    __syncthreads();
    TMARK0(1, b * p.tiles + tile, 3);

    float pcx = 0.f, pcy = 0.f, rpmin = INFINITY, obj = 0.f;
    const float xc = p24_anchor_centre(xs, st);
    const float yc = p24_anchor_centre(ys, st);
    if (active) {
        pcx = S.row[warp][0][lane];
        pcy = S.row[warp][1][lane];
#pragma unroll
        for (int c = 2; c < 26; ++c) rpmin = fminf(rpmin, S.row[warp][c][lane]);
        obj = S.row[warp][26][lane];
    }
    double objpart = active ? (double)p24_bce_logits(obj, 0.0f) : 0.0;
    const bool no_prune = (p.flags & P24_F_NO_PRUNE) != 0;
    const bool no_filter = (p.flags & P24_F_NO_FILTER) != 0;
    const bool tiny = !(rpmin >= 0.25f);  // a tiny (or NaN) predicted radius: outside the validated range of the bound

    // ---- which GTs can matter for this tile at all?  The box of the tile's anchor centres against every GT's reject disc
    // and centre window (-> s_near), the box of its predicted centres against every GT's far2 (-> s_farl): the per-anchor
    // loops below only visit those ------------------------------------------------------------------------------------
    int* s_near = reinterpret_cast<int*>(s_rec + (size_t)p.Lmax * GT_REC);
    int* s_farl = s_near + p.Lmax;
    if (n <= 32) {
        // few GTs: the boxes would cost more than they save
        if (tid < n) {
            s_near[tid] = tid;
            s_farl[tid] = tid;
        }
        if (tid == 0) {
            S.nnear = n;
            S.nfarl = n;
        }
        __syncthreads();
    } else {
        const float big = 3.0e38f;
        float ax0 = active ? xc : big, ax1 = active ? xc : -big, ay0 = active ? yc : big, ay1 = active ? yc : -big;
        const bool pok = active && pcx == pcx && pcy == pcy;
        float px0 = pok ? pcx : big, px1 = pok ? pcx : -big, py0 = pok ? pcy : big, py1 = pok ? pcy : -big;
        float stm = active ? st : 0.0f;
        ax0 = -warp_max(-ax0); ax1 = warp_max(ax1); ay0 = -warp_max(-ay0); ay1 = warp_max(ay1);
        px0 = -warp_max(-px0); px1 = warp_max(px1); py0 = -warp_max(-py0); py1 = warp_max(py1);
        stm = warp_max(stm);
        const bool anynan = __any_sync(0xffffffffu, active && !pok);
        if (lane == 0) {
            S.box[warp][0] = ax0; S.box[warp][1] = ax1; S.box[warp][2] = ay0; S.box[warp][3] = ay1;
            S.box[warp][4] = anynan ? -big : px0; S.box[warp][5] = anynan ? big : px1;
            S.box[warp][6] = anynan ? -big : py0; S.box[warp][7] = anynan ? big : py1;
            S.box[warp][8] = stm;
        }
        if (tid == 0) {
            S.nnear = 0;
            S.nfarl = 0;
        }
        __syncthreads();
        ax0 = S.box[0][0]; ax1 = S.box[0][1]; ay0 = S.box[0][2]; ay1 = S.box[0][3];
        px0 = S.box[0][4]; px1 = S.box[0][5]; py0 = S.box[0][6]; py1 = S.box[0][7];
        stm = S.box[0][8];
#pragma unroll
        for (int w = 1; w < P24_WARPS; ++w) {
            ax0 = fminf(ax0, S.box[w][0]); ax1 = fmaxf(ax1, S.box[w][1]); ay0 = fminf(ay0, S.box[w][2]); ay1 = fmaxf(ay1, S.box[w][3]);
            px0 = fminf(px0, S.box[w][4]); px1 = fmaxf(px1, S.box[w][5]); py0 = fminf(py0, S.box[w][6]); py1 = fmaxf(py1, S.box[w][7]);
            stm = fmaxf(stm, S.box[w][8]);
        }
        const float wr = 2.5f * stm * 1.001f + 1e-3f;  // (a little more than the window's half side)
        for (int g = tid; g < n; g += P24_THREADS) {
            const float* r = s_rec + g * GT_REC;
            const float cx = r[GT_CX], cy = r[GT_CY];
            // distance from the GT centre to the box of the anchor centres (0 inside), with a little slack
            const float ddx = fmaxf(fmaxf(ax0 - cx, cx - ax1), 0.0f), ddy = fmaxf(fmaxf(ay0 - cy, cy - ay1), 0.0f);
            const float dmin2 = fmaf(ddx, ddx, ddy * ddy) * 0.999f - 1e-3f;
            const bool nearb = no_prune || !(dmin2 > r[GT_RREJ2]) || (ddx < wr && ddy < wr) || !(cx == cx) || !(cy == cy);
            if (nearb) s_near[atomicAdd(&S.nnear, 1)] = g;
            // largest distance from the GT centre to the box of the predicted centres
            const float fx = fmaxf(fabsf(cx - px0), fabsf(cx - px1)), fy = fmaxf(fabsf(cy - py0), fabsf(cy - py1));
            const float dmax2 = fmaf(fx, fx, fy * fy) * 1.001f + 1e-3f;
            if (!(dmax2 < r[GT_FAR2])) s_farl[atomicAdd(&S.nfarl, 1)] = g;
        }
        __syncthreads();
    }
    const int nnear = S.nnear, nfarl = S.nfarl;

    // ---- one pass over the GTs that can matter: centre windows, the inscribed-disc accept and a bit mask (by position in
    // s_near) of the GTs whose reject radius the anchor is inside (the only ones that may need a polygon test) ----------
    bool cheap = false;
    unsigned near[4] = {0u, 0u, 0u, 0u};  // GTs 0..127; beyond that every GT is handled in place (see below)
    const float r25 = 2.5f * st + 1e-3f * st;  // conservative pre-filter radius of the window test
    const float4* s_rec4 = reinterpret_cast<const float4*>(s_rec);
    {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            unsigned m = 0u;
            const int ge = min(32, nnear - w * 32);
            for (int j = 0; j < ge; ++j) {
                const int g = s_near[w * 32 + j];
                const float4 h = s_rec4[g * (GT_REC / 4)];
                const float dx = h.x - xc, dy = h.y - yc;
                const float d2 = fmaf(dx, dx, dy * dy);
                cheap |= d2 < h.z;
                m |= (d2 <= h.w ? 1u : 0u) << j;
                if (fmaxf(fabsf(dx), fabsf(dy)) < r25 && p24_in_centre(h.x, h.y, xc, yc, st)) cheap = true;
            }
            const unsigned all = ge >= 32 ? 0xFFFFFFFFu : ((ge > 0 ? (1u << ge) : 1u) - 1u);
            near[w] = no_prune ? all : m;
        }
        for (int q = 128; q < nnear; ++q) {  // more than 128 GTs near the tile: windows and discs of the rest
            const int g = s_near[q];
            const float4 h = s_rec4[g * (GT_REC / 4)];
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
                const int g = s_near[w * 32 + __ffs(m) - 1];
                m &= m - 1;
                const int slot = atomicAdd(&S.nitems, 1);
                if (slot < ITEM_CAP) {
                    S.items[slot] = (unsigned)tid | ((unsigned)g << 8);
                } else if (!mine) {
                    const float* rec = s_rec + g * GT_REC;
                    mine = no_prune ? p24_in_polygon_exact(rec + GT_VX, rec + GT_VY, xc, yc)
                                    : p24_in_polygon(rec + GT_VX, rec + GT_VY, xc, yc);
                }
            }
        }
        for (int q = 128; q < nnear && !mine; ++q) {  // more than 128 GTs near the tile: test the rest in place
            const int g = s_near[q];
            const float4 h = s_rec4[g * (GT_REC / 4)];
            const float dx = h.x - xc, dy = h.y - yc;
            if (no_prune || fmaf(dx, dx, dy * dy) <= h.w) {
                const float* rec = s_rec + g * GT_REC;
                mine = no_prune ? p24_in_polygon_exact(rec + GT_VX, rec + GT_VY, xc, yc)
                                : p24_in_polygon(rec + GT_VX, rec + GT_VY, xc, yc);
            }
        }
    }
    if (mine) S.cand[tid] = 1;
    __syncthreads();
    TMARK0(1, b * p.tiles + tile, 4);
    TMARK0(1, b * p.tiles + tile, 9);
    {
        const int nitems = min(S.nitems, ITEM_CAP);
        for (int i = tid; i < nitems; i += P24_THREADS) {
            const unsigned it = S.items[i];
            const int al = it & 0xFF;
            if (((volatile int*)S.cand)[al]) continue;  // already a candidate through another GT
            const int g = it >> 8;
            const float* rec = s_rec + g * GT_REC;
            const int aa = tile * P24_THREADS + al;
            const float st2 = p.strides[aa];
            const float axc = p24_anchor_centre(p.x_shifts[aa], st2);
            const float ayc = p24_anchor_centre(p.y_shifts[aa], st2);
            const bool in = no_prune ? p24_in_polygon_exact(rec + GT_VX, rec + GT_VY, axc, ayc)
                                     : p24_in_polygon(rec + GT_VX, rec + GT_VY, axc, ayc);
            if (in) S.cand[al] = 1;
        }
    }
    __syncthreads();
    TMARK0(1, b * p.tiles + tile, 5);
    const bool cand = active && (n > 0) && (cheap || S.cand[tid]);
    if (tid == 0) S.nitems = 0;
    __syncthreads();
    S.cand[tid] = (cand ? 1 : 0) | (tiny ? 2 : 0);  // (read by far_pair, after the next barrier)

    // ---- the (GT, candidate) pairs whose PREDICTED centre lies beyond the GT's far2 (the only pairs whose value can reach
    // the GT's top 10): bounds into the GTs' lists, again through a work list ------------------------------------------
    TMARK0(1, b * p.tiles + tile, 2);
    TMARK0(1, b * p.tiles + tile, 8);
    if (cand && !no_filter && !tiny) {
        unsigned fm[4] = {0u, 0u, 0u, 0u};
        int cnt = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int ge = min(32, nfarl - w * 32);
            unsigned m = 0u;
            for (int j = 0; j < ge; ++j) {
                const int g = s_farl[w * 32 + j];
                const float4 h = s_rec4[g * (GT_REC / 4)];
                const float px = h.x - pcx, py = h.y - pcy;
                m |= (!(fmaf(px, px, py * py) < s_rec[g * GT_REC + GT_FAR2]) ? 1u : 0u) << j;
            }
            fm[w] = m;
            cnt += __popc(m);
        }
        int at = cnt ? atomicAdd(&S.nitems, cnt) : 0;  // one reservation for all of the anchor's pairs
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            unsigned m = fm[w];
            while (m) {
                const int g = s_farl[w * 32 + __ffs(m) - 1];
                m &= m - 1;
                if (at < ITEM_CAP) S.items[at] = (unsigned)tid | ((unsigned)g << 8);
                else far_pair(p, s_rec, S, b, tile, g, tid);
                ++at;
            }
        }
        for (int q = 128; q < nfarl; ++q) {  // more than 128 far GTs: the rest in place
            const int g = s_farl[q];
            const float* rec = s_rec + g * GT_REC;
            const float px = rec[GT_CX] - pcx, py = rec[GT_CY] - pcy;
            if (!(fmaf(px, px, py * py) < rec[GT_FAR2])) far_pair(p, s_rec, S, b, tile, g, tid);
        }
    } else if (cand && !no_filter) {
        // a tiny (or NaN) predicted radius: every pair of the anchor goes to the lists without a bound
        for (int g = 0; g < n; ++g) far_pair(p, s_rec, S, b, tile, g, tid);
    }
    // ---- per-anchor outputs, candidate bitmap and count, the all-anchor objectness term -------------------------
    const unsigned bal = __ballot_sync(0xffffffffu, cand);
    objpart = warp_sum_d(objpart);
    if (lane == 0) {
        p.cbits[((long long)b * p.tiles + tile) * P24_WARPS + warp] = bal;
        if (bal) atomicAdd(&p.ncand[b], __popc(bal));
        if (p.sums28)
            atomicAdd((unsigned long long*)&p.acc_fix[24], (unsigned long long)__double2ll_rn(objpart * FIX_SCALE_OBJ));
    }
    if (active) {
        // every anchor starts as background; k_tail overwrites the claimed ones
        const long long o = (long long)b * p.A + a;
        p.fg_mask[o] = 0;
        p.matched_gt[o] = -1;
        p.pred_iou[o] = 0.0f;
    }
    __syncthreads();
    TMARK0(1, b * p.tiles + tile, 6);
    {
        const int nitems = min(S.nitems, ITEM_CAP);
        for (int i = tid; i < nitems; i += P24_THREADS) {
            const unsigned it = S.items[i];
            far_pair(p, s_rec, S, b, tile, (int)(it >> 8), (int)(it & 0xFF));
        }
    }
}

// k_pass: persistent CTAs (one wave: the grid never exceeds what the device holds at once) drawing work items from
// ticket counters: first the seed items (their own counter), then the anchor tiles (the long items), then the
// (GT, level) centre-window items.  Launched as a programmatic dependent of k_prep: the first tile's rows are in flight
// before the CTA waits for the records.
__global__ void __launch_bounds__(P24_THREADS, 4) k_pass(const __grid_constant__ Params p) {
    extern __shared__ float4 s_dyn4[];   // [Lmax * GT_REC] floats: the image's records | the scratch of the seed items
    __shared__ AnchorShared S;
    __shared__ int s_item, s_seed;
    float* s_rec = reinterpret_cast<float*>(s_dyn4);
    const int n_anchor = p.B * p.tiles;
    const int tid = threadIdx.x;
    // the first tickets do not depend on k_prep
    if (tid == 0) s_item = (int)atomicAdd(&p.ticket[TK_ITEM], 1u);
    __syncthreads();
    int item = s_item;
    bool staged = false;
    if (item < n_anchor) {
        stage_rows(p, S, item / p.tiles, item % p.tiles);
        staged = true;
    }
    pdl_wait();  // the records, thresholds and window pairs come from k_prep / k_seed
    TMARK0(1, 6000 + blockIdx.x, 0);
    // ---- the anchor tiles first (the long items), image-major: an image's tiles (and its records) stay together in time --
    while (item < n_anchor) {
        __syncthreads();
        if (tid == 0) s_item = (int)atomicAdd(&p.ticket[TK_ITEM], 1u);  // the next one, in flight meanwhile
        TMARK0(1, item, 0);
        anchor_part(p, s_rec, S, item / p.tiles, item % p.tiles, staged);
        TMARK0(1, item, 7);
        staged = false;
        __syncthreads();
        item = s_item;
    }
    // ---- then the window chunks: 32 centre-window pairs of the batch's list, one per 8-lane group -----------------------
    {
        const int wtot = (int)__ldcg(&p.ticket[TK_WTOT]);
        const int n_wchunk = (wtot + P24_THREADS / 8 - 1) / (P24_THREADS / 8);
        __syncthreads();
        if (tid == 0) s_seed = (int)atomicAdd(&p.ticket[TK_WIN], 1u);
        __syncthreads();
        int wchunk = s_seed;
        while (wchunk < n_wchunk) {
            __syncthreads();
            if (tid == 0) s_seed = (int)atomicAdd(&p.ticket[TK_WIN], 1u);  // the next one, in flight meanwhile
            const int e = wchunk * (P24_THREADS / 8) + (tid >> 3);
            TMARK0(1, min(3000 + wchunk, 4095), 0);
            if (e < wtot) {
                const int2 pr = __ldcg(p.wlist + e);
                window_pair(p, pr.x, pr.y, group_mask());
            }
            TMARK0(1, min(3000 + wchunk, 4095), 7);
            __syncthreads();
            wchunk = s_seed;
        }
    }
    TMARK0(1, 6000 + blockIdx.x, 1);
    pdl_trigger();
}

// -------------------------------------------------------------------------------------------
// k_tail: one thread-block CLUSTER of TAIL_CL CTAs per image (the image's work spreads over TAIL_CL SMs; the phases are
// separated by the hardware cluster barrier instead of kernel boundaries)
// -------------------------------------------------------------------------------------------
#define TAIL_CL 8
#define TAIL_THREADS 256
#define TAIL_WARPS (TAIL_THREADS / 32)
#define TAIL_SURV 64   // survivors a GT warp evaluates per batch
#define ROW_PAD 108    // floats per staged row (27 + nc <= ROW_PAD is required for staging; else rows are read in place)
#define MBOX_SLOT 64   // floats per (epoch, rank) slot of a mailbox: 28 (value, epoch) words of 8 bytes
#define MBOX_EPOCHS 4  // slots alternate with the epoch: a rank runs at most two steps ahead of its own collect kernel
#define MBOX_FLAGS (MBOX_EPOCHS * P24_MAX_RANKS * MBOX_SLOT)  // own flags behind the slots: [1] epoch finished by k_fin

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

#define TAIL_WSEL 128  // entries of a GT's window table that can be valid (<= 25 per level)
struct TailShared {
    unsigned long long acc[26];   // fixed-point sums of the CTA (shared-memory atomics)
    KV kv[TAIL_WARPS];
    float sums[28];
    int surv[TAIL_WARPS][TAIL_SURV];      // per GT warp: anchors of the list entries that survive the refined threshold
    float sval[TAIL_WARPS][TAIL_SURV];    // ... their value bounds (upper; then the exact values)
    float slb[TAIL_WARPS][TAIL_SURV];     // ... and lower bounds
    float wc[TAIL_WARPS][TAIL_WSEL];      // per GT warp: the window-table entries that can be among the k cheapest
    int wa[TAIL_WARPS][TAIL_WSEL];
    int nuniq, last;
};

template <bool MAX>
__device__ __forceinline__ KV tail_block_select(KV x, KV* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    x = warp_select<MAX>(x);
    __syncthreads();
    if (lane == 0) s_red[warp] = x;
    __syncthreads();
    KV y = s_red[lane < TAIL_WARPS ? lane : 0];
    y = warp_select<MAX>(y);
    return y;
}

// sorted insert into a descending register list of P24_TOPK values
__device__ __forceinline__ void top_insert_desc(float (&t)[P24_TOPK], float v) {
#pragma unroll
    for (int q = 0; q < P24_TOPK; ++q) {
        if (v > t[q]) {
            const float x = t[q];
            t[q] = v;
            v = x;
        }
    }
}

// the `want` (<= 10) largest values held in the lanes' descending lists, popped in descending order by the whole warp:
// returns the r-th largest in every lane, one call per r
__device__ __forceinline__ float warp_pop_max(float (&t)[P24_TOPK]) {
    const int lane = threadIdx.x & 31;
    const KV best = warp_select<true>(KV{t[0], t[0] > P24_NEG_INF ? lane : 0x7fffffff});
    if (lane == best.i) {
#pragma unroll
        for (int q = 0; q < P24_TOPK - 1; ++q) t[q] = t[q + 1];
        t[P24_TOPK - 1] = P24_NEG_INF;
    }
    return best.i == 0x7fffffff ? P24_NEG_INF : best.v;
}

// Brute force (list overflow, P24_F_NO_FILTER, or fewer list entries than expected): the exact pair value of EVERY
// candidate of the image (candidate bitmap).  Every CTA of the cluster scans its share of the anchors and leaves its
// kc largest values (descending, -inf padded) in `out`; the first CTA merges them (brute_merge).  Whole CTA.
__device__ __noinline__ void brute_partial(const Params& p, TailShared& S, const float* __restrict__ rec, int b, int kc, int cr,
                                           float* __restrict__ out) {
    const int tid = threadIdx.x;
    const unsigned* bits = p.cbits + (long long)b * p.tiles * P24_WARPS;
    const int per = (p.A + TAIL_CL - 1) / TAIL_CL;
    float t[P24_TOPK];
#pragma unroll
    for (int q = 0; q < P24_TOPK; ++q) t[q] = P24_NEG_INF;
    for (int a = cr * per + tid; a < min(p.A, (cr + 1) * per); a += TAIL_THREADS) {
        if (!((__ldcg(bits + (a >> 5)) >> (a & 31)) & 1u)) continue;
        float v = pair_value_src(rec, src_of(p, b, a));
        if (!(v == v)) v = P24_POS_INF;  // NaN sorts first (torch.topk)
        top_insert_desc(t, v);
    }
    for (int r = 0; r < P24_TOPK; ++r) {
        KV best = {P24_NEG_INF, 0x7fffffff};
        if (r < kc) {
            best = tail_block_select<true>(KV{t[0], t[0] > P24_NEG_INF ? tid : 0x7fffffff}, S.kv);
            if (tid == best.i) {
#pragma unroll
                for (int q = 0; q < P24_TOPK - 1; ++q) t[q] = t[q + 1];
                t[P24_TOPK - 1] = P24_NEG_INF;
            }
        }
        if (tid == 0) out[r] = best.i == 0x7fffffff ? P24_NEG_INF : best.v;
    }
}

// the kc largest of the cluster's partial lists, summed in descending order; one warp
__device__ __forceinline__ float brute_merge(const float* __restrict__ parts, int kc) {
    const int lane = threadIdx.x & 31;
    float t[P24_TOPK];
#pragma unroll
    for (int q = 0; q < P24_TOPK; ++q) t[q] = P24_NEG_INF;
    for (int i = lane; i < TAIL_CL * P24_TOPK; i += 32) top_insert_desc(t, __ldcg(parts + i));
    float ksum = 0.0f;
    for (int r = 0; r < kc; ++r) {
        const float v = warp_pop_max(t);
        if (v == P24_NEG_INF) break;
        ksum = ksum + (v == P24_POS_INF ? NAN : v);
    }
    return ksum;
}

// Spill path (rare: GT with fewer valid anchors than its dynamic k): `need` more anchors with the smallest PENALISED
// cost among the candidates that are not valid for this GT (losses.py:460-464 on the penalised rows).  Ties -> lower
// anchor index.  Whole CTA; the new claims go to claim[at..] (global).
__device__ __noinline__ void spill_claims(const Params& p, TailShared& S, const float* __restrict__ rec, int b, int g, int need,
                                          int* claim, int at) {
    float lv[P24_TOPK];
    int li[P24_TOPK];
#pragma unroll
    for (int i = 0; i < P24_TOPK; ++i) {
        lv[i] = P24_POS_INF;
        li[i] = 0x7fffffff;
    }
    const int tid = threadIdx.x;
    const unsigned* bits = p.cbits + (long long)b * p.tiles * P24_WARPS;
    const float* tab = p.wtab + ((long long)b * p.Lmax + g) * P24_WT_STRIDE;
    const int c = gt_class(rec, p.nc);
    for (int a = tid; a < p.A; a += TAIL_THREADS) {
        if (!((__ldcg(bits + (a >> 5)) >> (a & 31)) & 1u)) continue;
        // valid for this GT (in window and in polygon: a finite entry of the window table)?
        int l = 0;
#pragma unroll
        for (int q = 1; q < P24_MAX_LEVELS; ++q) l += (q < p.nlev && a >= p.lev[q].off) ? 1 : 0;
        const int r = a - p.lev[l].off;
        const int iy = r / p.lev[l].W, ix = r - iy * p.lev[l].W;
        const int sx = ix - __float_as_int(__ldcg(tab + P24_WT_HDR + 2 * l));
        const int sy = iy - __float_as_int(__ldcg(tab + P24_WT_HDR + 2 * l + 1));
        if (sx >= 0 && sx < P24_WSIDE && sy >= 0 && sy < P24_WSIDE &&
            __ldcg(tab + l * P24_WSLOTS + sy * P24_WSIDE + sx) < P24_POS_INF)
            continue;
        const Src src = src_of(p, b, a);
        const float eo1 = 1.0f + expf(-src_obj(src));
        const float neg = thread_cls_neg_sum_src(src, p.nc, eo1);
        const float v = pair_value_src(rec, src);
        const float cost = p24_cost(cls_cost_from(neg, src_cls(src, c), 1.0f / eo1), v, false);
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
    int got = 0;
    for (int r = 0; r < need; ++r) {
        const KV head = {lv[0], li[0]};
        const KV win = tail_block_select<false>(head, S.kv);
        if (win.i == 0x7fffffff) break;  // fewer candidates than needed
        if (li[0] == win.i && lv[0] == win.v) {
            claim[at + got] = win.i;
#pragma unroll
            for (int q = 0; q < P24_TOPK - 1; ++q) {
                lv[q] = lv[q + 1];
                li[q] = li[q + 1];
            }
            lv[P24_TOPK - 1] = P24_POS_INF;
            li[P24_TOPK - 1] = 0x7fffffff;
        }
        ++got;
    }
}

// The k smallest costs among the GT's valid pairs (window table) -> claim[0 .. 10) of the GT (global; unused slots -1)
// (losses.py:460-464; ties -> lower anchor index).  One warp: the 10th smallest of the lanes' minima bounds the k-th
// smallest cost from above; the few entries up to it are ranked by counting.  Returns the number of valid pairs taken
// (< k: the GT must spill).
#define WSL_PER_LANE ((P24_WT_HDR + 31) / 32)   // 7
__device__ __forceinline__ int warp_select_claims(const Params& p, TailShared& S, int b, int g, int k, int* claim) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* tab = p.wtab + ((long long)b * p.Lmax + g) * P24_WT_STRIDE;
    const int nslot = P24_WSLOTS * p.nlev;
    float wc[WSL_PER_LANE];
    int nv = 0;
    // (costs and origins requested together: one round trip)
    const int org = lane < 2 * P24_MAX_LEVELS ? __float_as_int(__ldcg(tab + P24_WT_HDR + lane)) : 0;
    float lmin = P24_POS_INF;
#pragma unroll
    for (int q = 0; q < WSL_PER_LANE; ++q) {
        const int s = lane + 32 * q;
        float c = s < nslot ? __ldcg(tab + s) : P24_POS_INF;
        if (!(c < P24_POS_INF)) c = P24_POS_INF;  // (NaN cannot occur: the table holds finite costs or +inf)
        wc[q] = c;
        nv += c < P24_POS_INF ? 1 : 0;
        lmin = fminf(lmin, c);
    }
    nv = warp_sum_i(nv);
    const int take = min(k, nv);
    // at least 10 entries (or all valid ones) are <= tau: the k <= 10 cheapest are among the entries <= tau
    const int r0 = lane_rank<false>(lmin);
    const unsigned pick = __ballot_sync(0xffffffffu, r0 == P24_TOPK - 1);
    float tau = __shfl_sync(0xffffffffu, lmin, pick ? __ffs(pick) - 1 : 0);
    if (!pick) tau = P24_POS_INF;
    int ncomp = 0;
#pragma unroll
    for (int q = 0; q < WSL_PER_LANE; ++q) {
        const int s = lane + 32 * q;
        const bool in = wc[q] < P24_POS_INF && wc[q] <= tau;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int l = min(s / P24_WSLOTS, P24_MAX_LEVELS - 1), r = s - l * P24_WSLOTS;
        const int ox = __shfl_sync(0xffffffffu, org, 2 * l), oy = __shfl_sync(0xffffffffu, org, 2 * l + 1);
        if (in) {
            const int sy = r / P24_WSIDE, sx = r - sy * P24_WSIDE;
            const int at = ncomp + __popc(bal & ((1u << lane) - 1u));
            if (at < TAIL_WSEL) {
                S.wc[warp][at] = wc[q];
                S.wa[warp][at] = p.lev[l].off + (oy + sy) * p.lev[l].W + (ox + sx);
            }
        }
        ncomp += __popc(bal);
    }
    ncomp = min(ncomp, TAIL_WSEL);
    __syncwarp();
    if (lane < P24_TOPK) claim[lane] = -1;
    __syncwarp();
    for (int e = lane; e < ncomp; e += 32) {
        const float c = S.wc[warp][e];
        const int a = S.wa[warp][e];
        int rank = 0;
        for (int j = 0; j < ncomp; ++j) rank += kv_lt(S.wc[warp][j], S.wa[warp][j], c, a) ? 1 : 0;
        if (rank < take) claim[rank] = a;
    }
    __syncwarp();
    return take;
}

// Dynamic k of one GT from its list (one warp): the list holds (upper bound, anchor | all-apart flag) of every candidate
// pair whose bound reaches T.  (A) a certified lower bound of the 10th largest value refines the threshold: the 10th
// largest of the lanes' largest lower bounds (10 distinct entries reach it); (B) the entries whose bound still reaches it
// are evaluated exactly (8-lane groups, rows requested up front); the kc largest exact values are summed in descending
// order (like torch.topk(...).sum()).  Returns NaN-free sums only for clean inputs; `ok` = false: too many survivors
// for the buffers (the caller takes the brute-force path).
__device__ __forceinline__ float warp_topk_sum_list(const Params& p, TailShared& S, const float* __restrict__ rec, int b, int slot,
                                                    int lc, int kc, bool& ok) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float2* lst = p.list + (long long)slot * P24_LISTCAP;
    ok = true;
    float lmax = P24_NEG_INF;
    float2 e0[8];  // the first 256 entries stay in registers (most lists end there: one round trip for the whole list)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int i = 32 * u + lane;
        e0[u] = i < lc ? __ldcg(lst + i) : make_float2(P24_NEG_INF, 0.0f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        // bound = value + 2e-5 (+- 3e-6) for an all-apart pair: bound - 7e-5 is a certified lower bound
        if (__float_as_int(e0[u].y) & 0x80000000) lmax = fmaxf(lmax, e0[u].x - 7e-5f);
    }
    for (int i0 = 256; i0 < lc; i0 += 256) {  // eight independent loads per lane in flight
        float2 e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + 32 * u + lane;
            e[u] = i < lc ? __ldcg(lst + i) : make_float2(P24_NEG_INF, 0.0f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (__float_as_int(e[u].y) & 0x80000000) lmax = fmaxf(lmax, e[u].x - 7e-5f);
    }
    TMARK(2, 1024 + b * 64 + (slot - b * p.Lmax), 2);
    float tref = rec[GT_T];
    {
        const int r0 = lane_rank<true>(lmax);
        const unsigned pick = __ballot_sync(0xffffffffu, r0 == P24_TOPK - 1);
        const float v10 = __shfl_sync(0xffffffffu, lmax, pick ? __ffs(pick) - 1 : 0);
        if (pick) tref = fmaxf(tref, v10);
    }
    TMARK(2, 1024 + b * 64 + (slot - b * p.Lmax), 3);
    // the entries whose bound reaches `thr`, compacted (anchor, upper bound, lower bound); returns their number (the buffers
    // hold the first TAIL_SURV of them)
    auto compact = [&](float thr) {
        int ns = 0;
        for (int i0 = 0; i0 < lc; i0 += 256) {
            float2 e[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 32 * u + lane;
                e[u] = i0 == 0 ? e0[u] : (i < lc ? __ldcg(lst + i) : make_float2(P24_NEG_INF, 0.0f));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool keep = !(e[u].x < thr);
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                const int at = ns + __popc(bal & ((1u << lane) - 1u));
                if (keep && at < TAIL_SURV) {
                    const int bits = __float_as_int(e[u].y);
                    S.surv[warp][at] = bits & 0x7fffffff;
                    S.sval[warp][at] = e[u].x;
                    S.slb[warp][at] = (bits & 0x80000000) ? e[u].x - 7e-5f : 0.0f;  // (pair values are >= 0)
                }
                ns += __popc(bal);
            }
        }
        __syncwarp();
        return ns;
    };
    const unsigned gm = group_mask();
    const int grp = lane >> 3, sub = lane & 7;
    const float gcx = rec[GT_CX], gcy = rec[GT_CY];
    // exact values of the first `cnt` buffered survivors into S.sval, 16 at a time: every 8-lane group requests the rows
    // of its 4 pairs, then evaluates them
    auto exact_eval = [&](int cnt) {
        for (int j0 = 0; j0 < cnt; j0 += 16) {
            float rp[4][3], pc[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 4 * u + grp;
                const Src src = src_of(p, b, S.surv[warp][min(j, cnt - 1)]);
                pc[u][0] = src_geo(src, 0);
                pc[u][1] = src_geo(src, 1);
#pragma unroll
                for (int q = 0; q < 3; ++q) rp[u][q] = src_geo(src, 2 + sub * 3 + q);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 4 * u + grp;
                const float d = p24_centre_dist(gcx, gcy, pc[u][0], pc[u][1]);
                float sm = 0.0f;
#pragma unroll
                for (int q = 0; q < 3; ++q) sm = sm + p24_ray_loss(rec[GT_RG + sub * 3 + q], rp[u][q], d);  // (inlined: no call on this latency chain)
                sm = group_sum(sm, gm);
                float v = (sm / 24.0f) / 2.0f;
                if (!(v == v)) v = P24_POS_INF;  // NaN sorts first (torch.topk)
                if (sub == 0 && j < cnt) S.sval[warp][j] = v;
            }
        }
        __syncwarp();
    };
    int nsurv = compact(tref);
    if (nsurv > TAIL_SURV) {
        // more survivors than the buffers hold (a GT with many near-ties or few all-apart pairs): the exact values of the
        // buffered ones give a certified threshold (their kc-th largest), then the list is compacted again
        exact_eval(TAIL_SURV);
        float t3 = P24_NEG_INF;
        for (int j = lane; j < TAIL_SURV; j += 32) {
            const float vj = S.sval[warp][j];
            int rank = 0;
            for (int i = 0; i < TAIL_SURV; ++i) rank += kv_gt(S.sval[warp][i], i, vj, j) ? 1 : 0;
            if (rank == kc - 1 && vj < P24_POS_INF) t3 = vj;
        }
        t3 = warp_max(t3);
        __syncwarp();
        if (t3 > tref) nsurv = compact(t3);
    }
#ifdef P24_TIMING
    if (lane == 0) {
        g_tstamp[2][1024 + b * 64 + (slot - b * p.Lmax)][8] = nsurv;
        g_tstamp[2][1024 + b * 64 + (slot - b * p.Lmax)][9] = lc;
    }
#endif
    TMARK(2, 1024 + b * 64 + (slot - b * p.Lmax), 10);
    if (nsurv > TAIL_SURV) {
        ok = false;
        return 0.0f;
    }
    // ---- bracket: every survivor's value lies in [lower, upper]; the 10 largest lower bounds and the 10 largest upper
    // bounds all belong to survivors, so  L = sum of the 10 largest lower bounds <= sum of the 10 largest values <= U.
    // int(sum) is decided when floor(L) == floor(U): no exact evaluation (all-apart pairs: U - L = 7e-4) ---------------
    if (kc == P24_TOPK && nsurv >= P24_TOPK) {
        float ub[2], lb[2];
        int ru[2] = {0, 0}, rl[2] = {0, 0};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = lane + 32 * u;
            ub[u] = j < nsurv ? S.sval[warp][j] : P24_NEG_INF;
            lb[u] = j < nsurv ? S.slb[warp][j] : P24_NEG_INF;
        }
        for (int i = 0; i < nsurv; ++i) {
            const float ui = S.sval[warp][i], li = S.slb[warp][i];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                ru[u] += kv_gt(ui, i, ub[u], lane + 32 * u) ? 1 : 0;
                rl[u] += kv_gt(li, i, lb[u], lane + 32 * u) ? 1 : 0;
            }
        }
        float su = 0.0f, sl = 0.0f;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (lane + 32 * u < nsurv && ru[u] < P24_TOPK) su += ub[u];
            if (lane + 32 * u < nsurv && rl[u] < P24_TOPK) sl += lb[u];
        }
        su = warp_sum(su);
        sl = warp_sum(sl);
        const float fl = floorf(sl - 1e-4f), fu = floorf(su + 1e-4f);
        if (fl == fu && fl >= 1.0f) {
            TMARK(2, 1024 + b * 64 + (slot - b * p.Lmax), 4);
            return fl + 0.5f;  // any value with the decided integer part
        }
    }
    if (lane == 0) atomicAdd(&p.status[ST_EXACT], 1);
    exact_eval(nsurv);
    TMARK(2, 1024 + b * 64 + (slot - b * p.Lmax), 4);
    // the kc largest by rank counting, then summed in descending order by one lane
    float* top = reinterpret_cast<float*>(S.surv[warp]);  // (the anchors are no longer needed)
    float mine[2];
    int rk[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int j = lane + 32 * u;
        mine[u] = j < nsurv ? S.sval[warp][j] : P24_NEG_INF;
        rk[u] = 0;
    }
    for (int i = 0; i < nsurv; ++i) {
        const float vi = S.sval[warp][i];
#pragma unroll
        for (int u = 0; u < 2; ++u) rk[u] += kv_gt(vi, i, mine[u], lane + 32 * u) ? 1 : 0;
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 2; ++u)
        if (lane + 32 * u < nsurv && rk[u] < kc) top[rk[u]] = mine[u];
    __syncwarp();
    float ksum = 0.0f;
    const int have = min(kc, nsurv);
    for (int r = 0; r < have; ++r) {
        const float v = top[r];
        ksum = ksum + (v == P24_POS_INF ? NAN : v);
    }
    __syncwarp();
    return ksum;
}

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

__global__ void __launch_bounds__(TAIL_THREADS) k_tail(const __grid_constant__ Params p) {
    // recs [n * GT_REC] floats | org [n * 8] ints | claim [n * 10] ints | uniq [cap] ints | best [cap] u64 | rows [cap * ROW_PAD]
    extern __shared__ float4 t_dyn4[];
    __shared__ TailShared S;
    pdl_trigger();  // the next step's k_prep may become resident (it waits for this grid's completion)
    pdl_wait();
    const int b = blockIdx.x / TAIL_CL, cr = blockIdx.x % TAIL_CL, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int n = p.num_gt[b];
    const int ncand = __ldcg(&p.ncand[b]);
    const int kc = min(P24_TOPK, ncand);  // losses.py:452
    const int cap = (p.Lmax * P24_TOPK + TAIL_CL - 1) / TAIL_CL;  // claim slots (hence distinct anchors) per CTA
    float* s_rec = reinterpret_cast<float*>(t_dyn4);
    int* org = reinterpret_cast<int*>(s_rec + p.Lmax * GT_REC);
    int* claim = org + p.Lmax * 2 * P24_MAX_LEVELS;
    int* uniq = claim + p.Lmax * P24_TOPK;
    unsigned long long* best = reinterpret_cast<unsigned long long*>(uniq + ((cap + 1) & ~1));
    float* rows = reinterpret_cast<float*>(best + cap);
    const bool no_filter = (p.flags & P24_F_NO_FILTER) != 0;
    int* claimg = p.claimg + (long long)b * p.Lmax * P24_TOPK;
    int* kreq = p.kreq + b * p.Lmax;    // per GT: >= 0 clamped dynamic k a rare path must still honour, -1 none, -2 brute force
    int* ntake = p.ntake + b * p.Lmax;  // per GT: valid pairs already claimed

    {
        const float4* gsrc = reinterpret_cast<const float4*>(p.gt_rec + (long long)b * p.Lmax * GT_REC);
        for (int i = tid; i < n * (GT_REC / 4); i += TAIL_THREADS) t_dyn4[i] = __ldcg(gsrc + i);
        for (int i = tid; i < n * 2 * P24_MAX_LEVELS; i += TAIL_THREADS) {
            const int g = i / (2 * P24_MAX_LEVELS), q = i - g * (2 * P24_MAX_LEVELS);
            org[i] = __float_as_int(__ldcg(p.wtab + ((long long)b * p.Lmax + g) * P24_WT_STRIDE + P24_WT_HDR + q));
        }
    }
    if (tid == 0) S.nuniq = 0;
    if (tid < 26) S.acc[tid] = 0ull;
    TMARK0(2, blockIdx.x, 0);
    __syncthreads();
    TMARK0(2, blockIdx.x, 1);

    // ---- phase 1, one warp per GT (the image's GTs spread over the warps of the cluster): dynamic k from the GT's list,
    // then the k cheapest valid pairs -> claims (global) ---------------------------------------------------------------
    for (int g = warp * TAIL_CL + cr; g < n; g += TAIL_CL * TAIL_WARPS) {  // GT g -> CTA g % 8: spread over the cluster's SMs
        const int slot = b * p.Lmax + g;
        TMARK(2, 1024 + b * 64 + g, 0);
        const int lc = __ldcg(&p.lcount[slot]);
        TMARK(2, 1024 + b * 64 + g, 1);
        if (lane == 0) {
            p.lcount[slot] = 0;  // ready for the next call
            atomicMax(&p.status[ST_LISTMAX], lc);
            atomicAdd(&p.status[ST_LISTSUM], min(lc, 1 << 20));
            atomicAdd(&p.status[ST_GTS], 1);
        }
        int* cl = claimg + g * P24_TOPK;
        if (no_filter || lc > P24_LISTCAP || lc < kc) {
            if (lane == 0) {
                kreq[g] = -2;  // brute force, by the cluster's first CTA
                atomicAdd(&p.rare[b], 1);
            }
            if (lane < P24_TOPK) cl[lane] = -1;
            continue;
        }
        bool ok;
        const float ksum = warp_topk_sum_list(p, S, s_rec + g * GT_REC, b, slot, lc, kc, ok);
        TMARK(2, 1024 + b * 64 + g, 5);
        if (!ok) {
            if (lane == 0) {
                kreq[g] = -2;
                atomicAdd(&p.rare[b], 1);
            }
            if (lane < P24_TOPK) cl[lane] = -1;
            continue;
        }
        int k = (int)ksum;  // dynamic k = clamp(int(sum of the top-kc values), min=1)   losses.py:454-456
        if (k < 1) k = 1;
        const int kk = min(k, ncand);  // torch.topk would raise beyond the candidate count; clamp instead
        const int take = warp_select_claims(p, S, b, g, kk, cl);
        TMARK(2, 1024 + b * 64 + g, 6);
        if (lane == 0) {
            p.dyn_k[slot] = kk;
            ntake[g] = take;
            kreq[g] = take < kk ? kk : -1;
            if (take < kk) atomicAdd(&p.rare[b], 1);
        }
    }
    if (cr == 0)
        for (int g = n + tid; g < p.Lmax; g += TAIL_THREADS) p.dyn_k[b * p.Lmax + g] = 0;
    TMARK0(2, blockIdx.x, 2);
    __threadfence();
    cluster_sync_all();
    TMARK0(2, blockIdx.x, 3);
    // ---- rare paths, the cluster's first CTA, one GT at a time -----------------------------------------------------
    if (__ldcg(&p.rare[b]) > 0) {  // (the same value in every CTA of the cluster: written before the barrier)
        int nb = 0, ns = 0;
        for (int g = 0; g < n; ++g) {
            if (__ldcg(&kreq[g]) != -2) continue;  // (the same in every CTA)
            ++nb;
            // brute force over the candidate bitmap: every CTA its share of the anchors, the first CTA merges
            brute_partial(p, S, s_rec + g * GT_REC, b, kc, cr, p.brute + ((long long)b * TAIL_CL + cr) * P24_TOPK);
            __threadfence();
            cluster_sync_all();
            if (cr == 0 && warp == 0) {
                const float ksum = brute_merge(p.brute + (long long)b * TAIL_CL * P24_TOPK, kc);
                int k = (int)ksum;
                if (k < 1) k = 1;
                const int kk = min(k, ncand);
                const int take = warp_select_claims(p, S, b, g, kk, claimg + g * P24_TOPK);
                if (lane == 0) {
                    p.dyn_k[b * p.Lmax + g] = kk;
                    ntake[g] = take;
                    kreq[g] = take < kk ? kk : -1;
                }
            }
            __threadfence();
            cluster_sync_all();
        }
        if (cr == 0) {
            for (int g = 0; g < n; ++g) {
                const int kk = __ldcg(&kreq[g]);
                if (kk < 0) continue;
                ++ns;
                spill_claims(p, S, s_rec + g * GT_REC, b, g, kk - __ldcg(&ntake[g]), claimg + g * P24_TOPK, __ldcg(&ntake[g]));
                __syncthreads();
            }
            if (tid == 0) {
                if (nb) atomicAdd(&p.status[ST_BRUTE], nb);
                if (ns) atomicAdd(&p.status[ST_SPILL], ns);
            }
            __threadfence();
        }
        cluster_sync_all();
    }
    // ---- phase 2: every CTA looks at all claims of the image and owns a slice of the slots: the distinct claimed anchors
    // (first claim of every anchor) of its slice and whether several GTs claim them -------------------------------------
    TMARK0(2, blockIdx.x, 4);
    const int nslots = n * P24_TOPK;
    for (int t = tid; t < nslots; t += TAIL_THREADS) claim[t] = __ldcg(claimg + t);
    __syncthreads();
    {
        const unsigned gm = group_mask();
        const int sub = tid & 7;
        const int per = (nslots + TAIL_CL - 1) / TAIL_CL;  // the image's claim slots spread evenly over the cluster (<= cap)
        const int tend = min(nslots, (cr + 1) * per);
        for (int t0 = cr * per; t0 < tend; t0 += TAIL_THREADS / 8) {
            const int t = t0 + (tid >> 3);
            const int a = t < tend ? claim[t] : -1;
            int first = 1, multi = 0;
            if (a >= 0) {
                for (int j = sub; j < nslots; j += 8) {
                    if (j != t && claim[j] == a) {
                        multi = 1;
                        if (j < t) first = 0;
                    }
                }
            }
            multi = group_sum_i(multi, gm);
            first = group_sum_i(first, gm);
            if (a >= 0 && first == 8 && sub == 0) {
                const int e = atomicAdd(&S.nuniq, 1);
                uniq[e] = t | (multi ? 0x40000000 : 0);
                best[e] = ~0ull;
            }
        }
    }
    __syncthreads();
    const int nuniq = S.nuniq;
    TMARK0(2, blockIdx.x, 5);
    if (tid == 0 && nuniq) atomicAdd(&p.num_fg[b], nuniq);  // every claimed anchor ends up foreground (losses.py:479)

    // ---- phase 3a: anchors claimed by several GTs: argmin of the cost over ALL GTs (losses.py:471-476).  Valid pairs
    // always beat penalised ones and their costs are in the GTs' window tables: one thread per (anchor, GT), the anchor's
    // slot in the GT's table follows from its grid cell and the table's origin; packed (cost, GT) minimum in shared memory.
    // 3b (the same round trip): the rows of the CTA's anchors staged in shared memory, one warp per row ----------------
    for (int q = tid; q < nuniq * n; q += TAIL_THREADS) {
        const int e = q / n, g = q - e * n;
        const int u = uniq[e];
        if (!(u & 0x40000000)) continue;
        const int a = claim[u & 0x3FFFFFFF];
        int l = 0;
#pragma unroll
        for (int w = 1; w < P24_MAX_LEVELS; ++w) l += (w < p.nlev && a >= p.lev[w].off) ? 1 : 0;
        const int r = a - p.lev[l].off;
        const int iy = r / p.lev[l].W, ix = r - iy * p.lev[l].W;
        const int sx = ix - org[g * 2 * P24_MAX_LEVELS + 2 * l], sy = iy - org[g * 2 * P24_MAX_LEVELS + 2 * l + 1];
        if (sx >= 0 && sx < P24_WSIDE && sy >= 0 && sy < P24_WSIDE) {
            const float c = __ldcg(p.wtab + ((long long)b * p.Lmax + g) * P24_WT_STRIDE + l * P24_WSLOTS + sy * P24_WSIDE + sx);
            if (c < P24_POS_INF) atomicMin(&best[e], ((unsigned long long)p24_ordered(c) << 32) | (unsigned)g);
        }
    }
    const int C = 27 + p.nc;
    const bool stage = C <= ROW_PAD;
    if (stage) {
        for (int e = warp; e < nuniq; e += TAIL_WARPS) {
            const Src src = src_of(p, b, claim[uniq[e] & 0x3FFFFFFF]);
            float* dst = rows + e * ROW_PAD;
#pragma unroll
            for (int q = 0; q < (ROW_PAD + 31) / 32; ++q) {
                const int c = lane + 32 * q;
                if (c < C) dst[c] = c < 26 ? src_geo(src, c) : (c == 26 ? src_obj(src) : src_cls(src, c - 27));
            }
        }
    }
    __syncthreads();
    TMARK0(2, blockIdx.x, 6);

    // ---- phase 3c, one 8-lane group per claimed anchor: outputs and loss terms (3 rays and every 8th class per lane).
    // Contributions are accumulated as fixed-point integers: the sums do not depend on the order. -----------------------
    {
        const unsigned gm = group_mask();
        const int sub = tid & 7;
        long long acc_r[3] = {0, 0, 0};   // sum of loss24[:, 3 sub + q]
        long long acc_o = 0, acc_c = 0;   // -sum of obj logits at fg; cls BCE (group leaders)
        for (int e = tid >> 3; e < nuniq; e += TAIL_THREADS / 8) {
            const int u = uniq[e];
            const int t = u & 0x3FFFFFFF;
            const int aa = claim[t];
            int g = t / P24_TOPK;
            const float* row = stage ? rows + e * ROW_PAD : p.outputs + (long long)b * p.img_stride + (long long)aa * p.row_stride;
            if (u & 0x40000000) {
                const unsigned long long bb = best[e];
                if (bb != ~0ull) {
                    g = (int)(bb & 0xFFFFFFFFull);
                } else {
                    // without any valid pair (every claim came from a spill) the penalised costs decide: argmin over all
                    // GTs (losses.py:471-476), first index on ties.  Rare.
                    const float eo1 = 1.0f + expf(-row[26]);
                    const float neg = group_cls_neg_sum(row + 27, p.nc, eo1, gm);
                    KV bst = {P24_POS_INF, 0x7fffffff};
                    for (int gg = 0; gg < n; ++gg) {
                        const float* rc = s_rec + gg * GT_REC;
                        const float vv = group_pair_value(rc, row, gm);
                        const float c = p24_cost(cls_cost_from(neg, row[27 + gt_class(rc, p.nc)], 1.0f / eo1), vv, false);
                        if (kv_lt(c, gg, bst.v, bst.i)) {
                            bst.v = c;
                            bst.i = gg;
                        }
                    }
                    g = bst.i != 0x7fffffff ? bst.i : 0;
                }
            }
            const float* rec = s_rec + g * GT_REC;
            const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], row[0], row[1]);
            float l[3], sm = 0.0f;
#pragma unroll 1
            for (int q = 0; q < 3; ++q) {
                l[q] = p24_ray_loss(rec[GT_RG + sub * 3 + q], row[2 + sub * 3 + q], d);  // (inlined: no call on this latency chain)
                sm = sm + l[q];
            }
            sm = group_sum(sm, gm);
            const float v = (sm / 24.0f) / 2.0f;  // pair value == pred_ious_this_matching (losses.py:491)
            if (sub == 0) {
                const long long o = (long long)b * p.A + aa;
                p.fg_mask[o] = 1;
                p.matched_gt[o] = g;
                p.pred_iou[o] = v;
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) acc_r[q] += to_fix((double)l[q]);
            if (p.sums28) {
                // sum_j BCEWithLogits(x_j, t_j), t = v at the GT class and 0 elsewhere (losses.py:246-248, 298-302):
                // sum_j softplus(x_j) - x_c * v; the softplus sum as the log of a per-lane product (one log per lane;
                // the product of a lane's factors is restarted before it can overflow)
                const int c = gt_class(rec, p.nc);
                float prod = 1.0f, big = 0.0f;
                for (int j = sub; j < p.nc; j += 8) {
                    const float x = row[27 + j];
                    if (x < 8.0f) {
                        prod *= 1.0f + __expf(x);
                        if (prod > 1.0e30f) {
                            big += logf(prod);
                            prod = 1.0f;
                        }
                    } else {
                        big += x + log1pf(expf(-x));
                    }
                }
                big += logf(prod);
                big = group_sum(big, gm);
                if (sub == 0) {
                    acc_o += __double2ll_rn(-(double)row[26] * FIX_SCALE_OBJ);
                    acc_c += to_fix((double)big - (double)row[27 + c] * (double)v);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (acc_r[q] != 0) atomicAdd(&S.acc[sub * 3 + q], (unsigned long long)acc_r[q]);
        if (acc_o != 0) atomicAdd(&S.acc[24], (unsigned long long)acc_o);
        if (acc_c != 0) atomicAdd(&S.acc[25], (unsigned long long)acc_c);
    }
    TMARK0(2, blockIdx.x, 7);
    __syncthreads();
    TMARK0(2, blockIdx.x, 8);
    if (tid < 26 && p.sums28) {
        const unsigned long long t = S.acc[tid];
        if (t != 0) atomicAdd((unsigned long long*)&p.acc_fix[tid], t);
    } else if (tid == 26 && p.sums28 && nuniq) {
        atomicAdd((unsigned long long*)&p.acc_fix[26], (unsigned long long)nuniq);
    } else if (tid == 27 && p.sums28 && cr == 0) {
        atomicAdd((unsigned long long*)&p.acc_fix[27], (unsigned long long)n);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned done = atomicAdd(&p.ticket[TK_TAIL], 1u);
        S.last = (done == (unsigned)(p.B * TAIL_CL) - 1u) ? 1 : 0;
    }
    __syncthreads();
    TMARK0(2, blockIdx.x, 9);
    if (!S.last) return;
    __threadfence();
    // ---- last CTA of the grid: the batch sums (integer adds: exact, order independent), per-image counters reset, then
    // finalize or publish ---------------------------------------------------------------------------------------------
    if (tid < 28) {
        const long long t = __ldcg(&p.acc_fix[tid]);
        p.acc_fix[tid] = 0;  // ready for the next call
        S.sums[tid] = tid < 26 ? (float)((double)t / (tid == 24 ? FIX_SCALE_OBJ : FIX_SCALE)) : (float)t;
    }
    if (tid == 32) {
        p.ticket[TK_ITEM] = 0u;  // ready for the next call
        p.ticket[TK_LEFF] = 0u;
        p.ticket[TK_TAIL] = 0u;
        p.ticket[TK_SEED] = 0u;
        p.ticket[TK_WIN] = 0u;
        p.ticket[TK_WTOT] = 0u;
    }
    for (int i = tid; i < p.B; i += TAIL_THREADS) {
        p.seed_done[i] = 0;
        p.ncand[i] = 0;
        p.rare[i] = 0;
    }
    __syncthreads();
    if (!p.sums28) return;
    if (p.nranks > 1) {
        // ---- fused all-reduce, publish side: my 28 sums into everybody's mailbox (P2P stores over NVLink), a flag with the
        // call's epoch behind them.  MBOX_EPOCHS slot sets alternate with the epoch.  k_fin (p24_comm_finish) collects. ----
        const unsigned ep = p.epoch;
        const int half = (int)(ep % MBOX_EPOCHS) * P24_MAX_RANKS;
        // flow control of the slot sets: not before my own collect kernel of epoch ep - 2 has finished (then every peer has
        // consumed epoch ep - 4, whose slots are overwritten here: see p24.h)
        if (tid == 0) {
            volatile unsigned* fin = reinterpret_cast<volatile unsigned*>(p.mbox[p.rank] + MBOX_FLAGS + 1);
            while ((int)(ep - *fin) > 2) __nanosleep(64);
        }
        __syncthreads();
        // every sum travels as one 8-byte word (value, epoch): the word is written atomically, so the receiver needs no
        // flag behind the data and this CTA no system-wide fence (the way NCCL's low-latency protocol moves small messages)
        for (int q = tid; q < 28 * p.nranks; q += TAIL_THREADS) {
            const int r = q / 28, i = q - r * 28;
            const unsigned long long w = ((unsigned long long)ep << 32) | (unsigned long long)__float_as_uint(S.sums[i]);
            *reinterpret_cast<volatile unsigned long long*>(p.mbox[r] + (half + p.rank) * MBOX_SLOT + 2 * i) = w;
        }
        return;
    }
    if (tid < 28) p.sums28[tid] = S.sums[tid];
    if (p.state26 && warp == 0) finalize_warp(S.sums, p.state26, p.result54, p.weights27);
}

// k_fin (several GPUs): collect side of the fused all-reduce (p24_comm_finish).  One warp: wait for the flag of every rank
// in my own mailbox, add the contributions in rank order (the same bits on every rank), finalize.  It spins without a
// time-out, like a NCCL kernel: a wrong loss is worse than a hang that the framework's watchdog reports.  The wait is
// measured (status word) so that rank skew can be told from link latency.  The host side launches it on a side stream
// with NO stream dependency on the chain (an event between two steps would break their programmatic overlap): it waits on
// a flag that the chain's last CTA sets.  The next step's kernels do not depend on the global sums and run meanwhile.
struct FinParams {
    float* mbox;       // my own mailbox
    int nranks;
    unsigned epoch;
    float* sums28;
    float* state26;
    float* result54;
    float* weights27;
    int* status;       // (may be NULL)
};

__global__ void __launch_bounds__(32) k_fin(const FinParams p) {
    const int lane = threadIdx.x;
    __shared__ float s_sums[28];
    const unsigned ep = p.epoch;
    const int half = (int)(ep % MBOX_EPOCHS) * P24_MAX_RANKS;
    // launched on a side stream without any stream dependency on the chain: every (value, epoch) word of every rank --
    // this rank's own included -- is polled until it carries the epoch (a one-warp kernel: it cannot keep the chain from
    // running); the values are added in rank order
    const long long t0 = clock64();
    float t = 0.0f;
    if (lane < 28) {
        for (int r = 0; r < p.nranks; ++r) {
            volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(p.mbox + (half + r) * MBOX_SLOT + 2 * lane);
            unsigned long long v = *w;
            while ((unsigned)(v >> 32) != ep) {
                __nanosleep(32);
                v = *w;
            }
            t += __uint_as_float((unsigned)v);
        }
        s_sums[lane] = t;
        p.sums28[lane] = t;
    }
    __syncwarp();
    if (lane == 0 && p.status) p.status[ST_WAITCYC] = (int)min((long long)0x7fffffff, clock64() - t0);
    __syncwarp();
    if (p.state26) finalize_warp(s_sums, p.state26, p.result54, p.weights27);
    __threadfence();
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile unsigned*>(p.mbox + MBOX_FLAGS + 1) = ep;
}

__global__ void k_finalize(const float* __restrict__ sums28, float* __restrict__ state26, float* __restrict__ result54,
                           float* __restrict__ weights_n27) {
    finalize_warp(sums28, state26, result54, weights_n27);
}

size_t pass_smem(int Lmax) { return (size_t)Lmax * GT_REC * sizeof(float) + (size_t)2 * Lmax * sizeof(int); }
size_t tail_smem(int Lmax, int nc) {
    const size_t cap = ((size_t)Lmax * P24_TOPK + TAIL_CL - 1) / TAIL_CL;
    size_t b = (size_t)Lmax * GT_REC * sizeof(float);                 // recs
    b += (size_t)Lmax * 2 * P24_MAX_LEVELS * sizeof(int);             // window origins
    b += (size_t)Lmax * P24_TOPK * sizeof(int);                       // claims
    b += ((cap + 1) & ~(size_t)1) * sizeof(int);                      // uniq
    b += cap * sizeof(unsigned long long);                            // best
    if (27 + nc <= ROW_PAD) b += cap * ROW_PAD * sizeof(float);       // staged rows
    return b;
}

template <typename K>
cudaError_t launch(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, const Params& p, int cluster = 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = (unsigned)cluster;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

template <typename K>
cudaError_t launch2(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, const Params& p, int arg) {
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
    return cudaLaunchKernelEx(&cfg, kernel, p, arg);
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

extern "C" int p24_read_status(void* workspace, int B, int A, int Lmax, int32_t* h_status8, void* stream) {
    if (!workspace || !h_status8 || B <= 0 || A <= 0 || Lmax <= 0) return P24_E_BADARG;
    const P24Workspace L = p24_layout(B, A, Lmax);
    char* st = (char*)workspace + L.status;
    cudaError_t e = cudaMemcpyAsync(h_status8, st, ST_WORDS * sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    // the counters restart with every read; the error bits (word 0) are sticky
    e = cudaMemsetAsync(st + sizeof(int), 0, (ST_WORDS - 1) * sizeof(int), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize((cudaStream_t)stream);
}

namespace {
int simota_impl(const float* outputs, int64_t img_stride, int64_t row_stride, const float* const* h_raw,
                const int64_t* h_raw_bs, int B, int A,
                                     int num_classes, const float* labels, int64_t lab_img_stride,
                                     int64_t lab_row_stride, int Lmax, const float* x_shifts, const float* y_shifts,
                                     const float* strides, const int32_t* h_levels, int n_levels, uint8_t* fg_mask,
                                     int32_t* matched_gt, float* pred_iou, int32_t* num_fg, int32_t* num_gt,
                                     int32_t* dyn_k, float* sums28, float* state26, float* result54,
                                     float* weights_n27, void* workspace, size_t workspace_bytes, uint32_t flags,
                                     void* const* h_mailboxes, int rank, int nranks, uint32_t epoch, void* stream) {
    if ((!outputs && !h_raw) || !labels || !x_shifts || !y_shifts || !strides || !fg_mask || !matched_gt || !pred_iou || !num_fg ||
        !num_gt || !dyn_k || !workspace || !h_levels)
        return P24_E_BADARG;
    if (h_raw && (!h_raw_bs || 27 + num_classes > ROW_PAD)) return h_raw_bs ? P24_E_UNSUPPORTED : P24_E_BADARG;
    if (B <= 0 || A <= 0 || Lmax <= 0 || num_classes <= 0 || Lmax > 65535 || B > 65535) return P24_E_BADARG;
    if (state26 && (!sums28 || !result54 || !weights_n27)) return P24_E_BADARG;
    if (n_levels <= 0) return P24_E_BADARG;
    if (n_levels > P24_MAX_LEVELS) return P24_E_UNSUPPORTED;
    const P24Workspace L = p24_layout(B, A, Lmax);
    if (workspace_bytes < L.total) return P24_E_WORKSPACE;
    if (((uintptr_t)workspace & 255) != 0) return P24_E_BADARG;
    const size_t dyn_pass = pass_smem(Lmax), dyn_tail = tail_smem(Lmax, num_classes);
    if (dyn_pass > 160 * 1024 || dyn_tail > 200 * 1024 || p24_tiles(A) > 65535) return P24_E_UNSUPPORTED;
    char* ws = (char*)workspace;
    Params p;
    p.outputs = h_raw ? nullptr : outputs; p.img_stride = img_stride; p.row_stride = row_stride;
    for (int t = 0; t < 3; ++t)
        for (int l = 0; l < P24_MAX_LEVELS; ++l) {
            p.raw[t][l] = nullptr;
            p.raw_bs[t][l] = 0;
            if (h_raw && l < n_levels && n_levels <= P24_MAX_LEVELS) {
                if (!h_raw[t * n_levels + l]) return P24_E_BADARG;
                p.raw[t][l] = h_raw[t * n_levels + l];
                p.raw_bs[t][l] = h_raw_bs[t * n_levels + l];
            }
        }
    p.B = B; p.A = A; p.nc = num_classes;
    p.labels = labels; p.lab_img_stride = lab_img_stride; p.lab_row_stride = lab_row_stride; p.Lmax = Lmax;
    p.x_shifts = x_shifts; p.y_shifts = y_shifts; p.strides = strides;
    p.fg_mask = fg_mask; p.matched_gt = matched_gt; p.pred_iou = pred_iou;
    p.num_fg = num_fg; p.num_gt = num_gt; p.dyn_k = dyn_k; p.sums28 = sums28;
    p.state26 = state26; p.result54 = result54; p.weights27 = weights_n27;
    p.ticket = (unsigned*)(ws + L.ticket);
    p.acc_fix = (long long*)(ws + L.acc_fix);
    p.status = (int*)(ws + L.status);
    p.seed_done = (int*)(ws + L.seed_done);
    p.ncand = (int*)(ws + L.ncand);
    p.lcount = (int*)(ws + L.lcount);
    p.gt_rec = (float*)(ws + L.gt_rec);
    p.wtab = (float*)(ws + L.wtab);
    p.list = (float2*)(ws + L.list);
    p.cbits = (unsigned*)(ws + L.cbits);
    p.wlist = (int2*)(ws + L.wlist);
    p.brute = (float*)(ws + L.brute);
    p.claimg = (int*)(ws + L.claimg);
    p.kreq = (int*)(ws + L.kreq);
    p.ntake = (int*)(ws + L.ntake);
    p.rare = (int*)(ws + L.rare);
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
            p.lev[l].off = 0; p.lev[l].W = 1; p.lev[l].H = 0; p.lev[l].st = 1.0f;
            if (l < n_levels) {
                const int32_t* d = h_levels + 4 * l;
                float stv;
                memcpy(&stv, d + 3, sizeof(float));
                if (d[0] != next || d[1] <= 0 || d[2] <= 0 || !(stv > 0.0f)) return P24_E_BADARG;
                p.lev[l].off = d[0]; p.lev[l].W = d[1]; p.lev[l].H = d[2]; p.lev[l].st = stv;
                next += (long long)d[1] * d[2];
            }
        }
        if (next != A) return P24_E_BADARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (p24::dev_once(1u << 0)) {  // per device: a process may drive several GPUs
        cudaFuncSetAttribute(k_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    const bool pdl = !(flags & P24_F_NO_PDL) && !p24::prof_on();
    cudaError_t e = cudaSuccess;
    const int n_sm = p24::dev_info().n_sm;
    p24::prof_mark(0, st);
    {
        // k_prep waits inside for the records of all CTAs of an image: one launch when the whole grid is resident at once
        // (always, at training batch sizes), else records and seeds in two launches
        const dim3 pgrid((Lmax + PREP_GTS - 1) / PREP_GTS, B);
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_prep, PREP_THREADS, dyn_pass);
        if (e != cudaSuccess) return (int)e;
        const bool fused = (long long)pgrid.x * pgrid.y <= (long long)per_sm * n_sm;
        e = launch2(k_prep, pgrid, dim3(PREP_THREADS), dyn_pass, st, pdl, p, fused ? 0 : 1);
        if (e != cudaSuccess) return (int)e;
        if (!fused) {
            e = launch2(k_prep, pgrid, dim3(PREP_THREADS), dyn_pass, st, pdl, p, 2);
            if (e != cudaSuccess) return (int)e;
        }
    }
    p24::prof_mark(1, st);
    {
        // one wave of persistent CTAs drawing tiles and window chunks from ticket counters
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pass, P24_THREADS, dyn_pass);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) return P24_E_UNSUPPORTED;
        const long long items = (long long)B * p.tiles > (long long)B * Lmax ? (long long)B * p.tiles : (long long)B * Lmax;
        const long long cap = (long long)per_sm * n_sm;
        e = launch(k_pass, dim3((unsigned)(items < cap ? items : cap)), dim3(P24_THREADS), dyn_pass, st, pdl, p);
        if (e != cudaSuccess) return (int)e;
    }
    p24::prof_mark(2, st);
    e = launch(k_tail, dim3(B * TAIL_CL), dim3(TAIL_THREADS), dyn_tail, st, pdl, p, TAIL_CL);
    if (e != cudaSuccess) return (int)e;
    p24::prof_mark(3, st);
    return (int)cudaGetLastError();
}
}  // namespace

extern "C" int p24_simota_loss_batch(const float* outputs, int64_t img_stride, int64_t row_stride, int B, int A,
                                     int num_classes, const float* labels, int64_t lab_img_stride,
                                     int64_t lab_row_stride, int Lmax, const float* x_shifts, const float* y_shifts,
                                     const float* strides, const int32_t* h_levels, int n_levels, uint8_t* fg_mask,
                                     int32_t* matched_gt, float* pred_iou, int32_t* num_fg, int32_t* num_gt,
                                     int32_t* dyn_k, float* sums28, float* state26, float* result54,
                                     float* weights_n27, void* workspace, size_t workspace_bytes, uint32_t flags,
                                     void* const* h_mailboxes, int rank, int nranks, uint32_t epoch, void* stream) {
    if (!outputs) return P24_E_BADARG;
    return simota_impl(outputs, img_stride, row_stride, nullptr, nullptr, B, A, num_classes, labels, lab_img_stride,
                       lab_row_stride, Lmax, x_shifts, y_shifts, strides, h_levels, n_levels, fg_mask, matched_gt, pred_iou,
                       num_fg, num_gt, dyn_k, sums28, state26, result54, weights_n27, workspace, workspace_bytes, flags,
                       h_mailboxes, rank, nranks, epoch, stream);
}

extern "C" int p24_simota_loss_batch_raw(const float* const* h_raw, const int64_t* h_raw_batch_stride, int B, int A,
                                         int num_classes, const float* labels, int64_t lab_img_stride,
                                         int64_t lab_row_stride, int Lmax, const float* x_shifts, const float* y_shifts,
                                         const float* strides, const int32_t* h_levels, int n_levels, uint8_t* fg_mask,
                                         int32_t* matched_gt, float* pred_iou, int32_t* num_fg, int32_t* num_gt,
                                         int32_t* dyn_k, float* sums28, float* state26, float* result54,
                                         float* weights_n27, void* workspace, size_t workspace_bytes, uint32_t flags,
                                         void* const* h_mailboxes, int rank, int nranks, uint32_t epoch, void* stream) {
    if (!h_raw || !h_raw_batch_stride) return P24_E_BADARG;
    return simota_impl(nullptr, 0, 0, h_raw, h_raw_batch_stride, B, A, num_classes, labels, lab_img_stride, lab_row_stride,
                       Lmax, x_shifts, y_shifts, strides, h_levels, n_levels, fg_mask, matched_gt, pred_iou, num_fg, num_gt,
                       dyn_k, sums28, state26, result54, weights_n27, workspace, workspace_bytes, flags, h_mailboxes, rank,
                       nranks, epoch, stream);
}

extern "C" size_t p24_comm_mailbox_bytes(void) { return (size_t)(MBOX_FLAGS + 64) * sizeof(float); }

extern "C" int p24_comm_finish(void* d_own_mailbox, int nranks, uint32_t epoch, float* sums28, float* state26, float* result54,
                               float* weights_n27, void* workspace, int B, int A, int Lmax, void* stream) {
    if (!d_own_mailbox || !sums28 || nranks < 2 || nranks > P24_MAX_RANKS) return P24_E_BADARG;
    if (state26 && (!result54 || !weights_n27)) return P24_E_BADARG;
    FinParams f;
    f.mbox = (float*)d_own_mailbox;
    f.nranks = nranks;
    f.epoch = epoch;
    f.sums28 = sums28; f.state26 = state26; f.result54 = result54; f.weights27 = weights_n27;
    f.status = nullptr;
    if (workspace && B > 0 && A > 0 && Lmax > 0) f.status = (int*)((char*)workspace + p24_layout(B, A, Lmax).status);
    k_fin<<<1, 32, 0, (cudaStream_t)stream>>>(f);
    return (int)cudaGetLastError();
}

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
    return (int)cudaMemcpyFromSymbol(h_out, g_tstamp, sizeof(unsigned long long) * 3 * TM_ROWS * TM_SLOTS);
}
#endif
