// p24_simota.cu — the fused YOLOX-24p SimOTA assignment + loss-sum path for sm_100a.
//
// Replaces, for a whole batch and without ever writing the G x A cost matrix to memory:
//   Loss_Function.get_assignments / get_in_boxes_info / pts_in_poly   models/losses.py:359-592
//   utils.boxes.bboxes_iou + pairwise circle_inter                   utils/boxes.py:102-243
//   Loss_Function.dynamic_k_matching                                 models/losses.py:444-494
//   the loss sums and re-weighting of Loss_Function.forward          models/losses.py:246-345
//
// Kernel chain (all on the caller's stream, no host synchronisation):
//   k_anchor_pass  one CTA per 256-anchor tile: the ONE pass over the head output (coalesced reads of the
//                  27 geometry channels of every row).  Builds the image's GT records, the candidate mask
//                  (polygon test OR centre window) with geometric pruning and an atan2-free angle test,
//                  the per-GT centre-window lists, the compacted candidate list and sum BCEWithLogits(obj, 0)
//   k_gt_match     one CTA per GT: dynamic k from the top-10-largest pair values over the candidates
//                  (bracketed by exact seed values and a monotone upper bound; exact evaluation of the few
//                  pairs the bound cannot exclude otherwise); exact cost of the GT's valid (in polygon AND in
//                  window) pairs, the k smallest -> claims (spill into the penalised regime when too few)
//   k_resolve_loss one CTA per tile: conflict resolution (argmin over all GTs), fg_mask / matched_gt /
//                  pred_iou, the 28 loss sums; the last CTA reduces them in a fixed order and, when asked,
//                  applies the normalisation and the stateful re-weighting (losses.py:280-345)
//
// Compile with -fmad=false: the discrete decisions hang on fp32 thresholds evaluated in the
// reference's operation order; bounds and fast paths use explicit fmaf.
#include "p24_common.cuh"

namespace {

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
    int* ccount;
    int* wcount;
    int* wlist;
    int* claim_cnt;
    int* claim_gt;
    double* obj_part;
    double* loss_part;
    unsigned* ticket;
    int* err_flag;
    unsigned flags;
    int tiles;
};

// -------------------------------------------------------------------------------------------
// GT records of one image, built by every CTA of the anchor pass into shared memory
// -------------------------------------------------------------------------------------------
// nlabel = (labels.sum(2) > 0).sum(1)   losses.py:190 ; the first n rows are the GTs (losses.py:219-220)
__device__ int count_labels(const Params& p, const float* lab, int* s_tmp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (p.flags & P24_F_ALL_ROWS) return p.Lmax;
    if (tid == 0) *s_tmp = 0;
    __syncthreads();
    int local = 0;
    for (int r = warp; r < p.Lmax; r += P24_WARPS) {
        const float* row = lab + (long long)r * p.lab_row_stride;
        double s = (double)row[lane];
        if (lane + 32 < 51) s += (double)row[lane + 32];
        s = warp_sum_d(s);
        if ((float)s > 0.0f) ++local;
    }
    if (lane == 0 && local) atomicAdd(s_tmp, local);
    __syncthreads();
    return *s_tmp;
}

__device__ void build_gt_record(const float* __restrict__ row, float* __restrict__ rec) {
    const float cx = row[1], cy = row[2];
    float rgmax = 0.0f, rgmin = INFINITY, perim = 0.0f, rin = INFINITY;
    bool inside = false;
    float x = row[3], y = row[4];
    const float x0 = x, y0 = y;
#pragma unroll 4
    for (int k = 0; k < P24_RAYS; ++k) {
        const float x2 = (k == P24_RAYS - 1) ? x0 : row[5 + 2 * k];
        const float y2 = (k == P24_RAYS - 1) ? y0 : row[6 + 2 * k];
        rec[GT_VX + k] = x;
        rec[GT_VY + k] = y;
        const float rg = p24_gt_radius(x - cx, y - cy);
        rec[GT_RG + k] = rg;
        rgmax = fmaxf(rgmax, rg);
        rgmin = fminf(rgmin, rg);
        const float ex = x2 - x, ey = y2 - y;
        const float len2 = fmaf(ex, ex, ey * ey);
        perim += sqrtf(len2);
        // distance from the centre to the edge segment
        const float wx = cx - x, wy = cy - y;
        float tt = len2 > 0.0f ? __fdividef(fmaf(wx, ex, wy * ey), len2) : 0.0f;
        tt = fminf(fmaxf(tt, 0.0f), 1.0f);
        const float qx = wx - tt * ex, qy = wy - tt * ey;
        rin = fminf(rin, sqrtf(fmaf(qx, qx, qy * qy)));
        // crossing-number parity of the centre
        if ((y > cy) != (y2 > cy)) {
            const float xi = fmaf(ex, __fdividef(cy - y, ey), x);
            if (cx < xi) inside = !inside;
        }
        x = x2;
        y = y2;
    }
    // A point inside a closed polygon has |winding| >= 1, so its total unsigned angle is >= 360 degrees: a disc
    // around an interior centre that stays clear of every edge passes the >= 350 test (2 % + 0.01 px of slack
    // covers the fp32 evaluation of the distances above).
    float ra = (inside && rin == rin) ? fmaf(0.98f, rin, -0.01f) : 0.0f;
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
}

// -------------------------------------------------------------------------------------------
// shared device helpers
// -------------------------------------------------------------------------------------------
// exact pair value of (GT record, prediction row in global memory): utils/boxes.py:166-243, one thread
__device__ float pair_value_row(const float* __restrict__ rec, const float* __restrict__ row) {
    const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], row[0], row[1]);
    float s = 0.0f;
#pragma unroll 4
    for (int k = 0; k < P24_RAYS; ++k) s = s + p24_ray_loss(rec[GT_RG + k], row[2 + k], d);
    return (s / 24.0f) / 2.0f;
}

__device__ __forceinline__ int gt_class(const float* rec, int nc) {
    const int c = (int)rec[GT_CLS];
    return min(max(c, 0), nc - 1);
}

// Warp-cooperative sum over all classes of BCE(p_j, 0) (losses.py:406-416) in product form; every lane returns it.
__device__ float warp_cls_neg_sum(const float* __restrict__ cls, int nc, float eo1) {
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

// single-thread version of the same sum (rare slow paths)
__device__ float thread_cls_neg_sum(const float* __restrict__ cls, int nc, float eo1) {
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
// k_anchor_pass
// -------------------------------------------------------------------------------------------
#define ITEM_CAP 2048
#define G_CHUNK 8
#define ROW_CH 27  // channels 0..26 of a head row are read here: centre, 24 radii, objectness

__global__ void __launch_bounds__(P24_THREADS) k_anchor_pass(Params p) {
    extern __shared__ float4 s_dyn4[];
    float* s_gt = reinterpret_cast<float*>(s_dyn4);  // [n * GT_REC]
    const int b = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int a = tile * P24_THREADS + tid;
    const bool active = a < p.A;

    __shared__ float s_row[P24_WARPS][ROW_CH][33];
    __shared__ unsigned s_items[ITEM_CAP];
    __shared__ int s_cand[P24_THREADS];
    __shared__ int s_nitems, s_tmp;
    __shared__ int s_wcnt[P24_WARPS];
    __shared__ double s_red[P24_WARPS];

    // ---- GT records -----------------------------------------------------------------------------
    const float* lab = p.labels + (long long)b * p.lab_img_stride;
    const int n = count_labels(p, lab, &s_tmp);
    for (int g = tid; g < n; g += P24_THREADS) build_gt_record(lab + (long long)g * p.lab_row_stride, s_gt + g * GT_REC);
    s_cand[tid] = 0;

    // ---- coalesced load of the tile's rows: each warp reads its 32 rows, 27 contiguous floats at a time --
    const float* img = p.outputs + (long long)b * p.img_stride;
    {
        const int a0 = tile * P24_THREADS + warp * 32;
        const int nrow = min(32, p.A - a0);
        if (lane < ROW_CH) {
#pragma unroll 8
            for (int r = 0; r < nrow; ++r) s_row[warp][lane][r] = img[(long long)(a0 + r) * p.row_stride + lane];
        }
    }
    __syncthreads();
    if (tile == 0) {
        float* gdst = p.gt_rec + (long long)b * p.Lmax * GT_REC;
        for (int i = tid; i < n * GT_REC; i += P24_THREADS) gdst[i] = s_gt[i];
        if (tid == 0) {
            p.num_gt[b] = n;
            p.num_fg[b] = 0;
        }
    }
    float pcx = 0.f, pcy = 0.f, rpmax = 0.f, rpmin = INFINITY, obj = 0.f;
    float xc = 0.f, yc = 0.f, st = 1.f;
    if (active) {
        pcx = s_row[warp][0][lane];
        pcy = s_row[warp][1][lane];
#pragma unroll
        for (int c = 2; c < 26; ++c) {
            const float v = s_row[warp][c][lane];
            rpmax = fmaxf(rpmax, v);
            rpmin = fminf(rpmin, v);
        }
        obj = s_row[warp][26][lane];
        st = p.strides[a];
        xc = p24_anchor_centre(p.x_shifts[a], st);
        yc = p24_anchor_centre(p.y_shifts[a], st);
    }
    double objpart = active ? (double)p24_bce_logits(obj, 0.0f) : 0.0;

    // ---- pass 1 over the GTs: centre windows (-> per-GT lists) and the inscribed-disc accept ------------
    bool cheap = false;
    if (active) {
        const float r25 = 2.5f * st + 1e-3f * st;  // conservative pre-filter radius of the window test
        for (int g = 0; g < n; ++g) {
            const float4 h = s_dyn4[g * (GT_REC / 4)];
            const float dx = h.x - xc, dy = h.y - yc;
            const float d2 = fmaf(dx, dx, dy * dy);
            cheap |= d2 < h.z;
            if (fmaxf(fabsf(dx), fabsf(dy)) < r25 && p24_in_centre(h.x, h.y, xc, yc, st)) {
                cheap = true;
                const int slot = atomicAdd(&p.wcount[b * p.Lmax + g], 1);
                if (slot < P24_VCAP) p.wlist[((long long)b * p.Lmax + g) * P24_VCAP + slot] = a;
                else atomicOr(p.err_flag, 1);
            }
        }
    }
    const bool no_prune = (p.flags & P24_F_NO_PRUNE) != 0;

    // ---- pass 2: anchors not yet accepted need a polygon test against every GT whose reject radius they
    // are inside; the tests are compacted into a work list so that all threads stay busy -----------------
    for (int g0 = 0; g0 < n; g0 += G_CHUNK) {
        if (tid == 0) s_nitems = 0;
        __syncthreads();
        if (active && (!cheap || no_prune)) {
            const int g1 = min(g0 + G_CHUNK, n);
            for (int g = g0; g < g1; ++g) {
                const float4 h = s_dyn4[g * (GT_REC / 4)];
                const float dx = h.x - xc, dy = h.y - yc;
                const float d2 = fmaf(dx, dx, dy * dy);
                if (no_prune || (d2 <= h.w && !(d2 < h.z)))
                    s_items[atomicAdd(&s_nitems, 1)] = (unsigned)tid | ((unsigned)g << 8);
            }
        }
        __syncthreads();
        const int nitems = s_nitems;
        for (int i = tid; i < nitems; i += P24_THREADS) {
            const unsigned it = s_items[i];
            const int al = it & 0xFF;
            if (((volatile int*)s_cand)[al]) continue;  // already a candidate through another GT
            const int g = it >> 8;
            const float* rec = s_gt + g * GT_REC;
            const int aa = tile * P24_THREADS + al;
            const float st2 = p.strides[aa];
            const float axc = p24_anchor_centre(p.x_shifts[aa], st2);
            const float ayc = p24_anchor_centre(p.y_shifts[aa], st2);
            const bool in = no_prune ? (p24_angle_sum(rec + GT_VX, rec + GT_VY, axc, ayc) >= 350.0f)
                                     : p24_in_polygon(rec + GT_VX, rec + GT_VY, axc, ayc);
            if (in) s_cand[al] = 1;
        }
        __syncthreads();
    }

    // ---- compacted candidate list of the tile (deterministic order) + per-anchor scratch reset ----------
    const bool cand = active && (n > 0) && (cheap || s_cand[tid]);
    const unsigned bal = __ballot_sync(0xffffffffu, cand);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < P24_WARPS; ++w) {
        const int c = s_wcnt[w];
        base += (w < warp) ? c : 0;
        total += c;
    }
    const long long blk = (long long)b * p.tiles + tile;
    if (cand) {
        const int rank = base + __popc(bal & ((1u << lane) - 1u));
        // a prediction with a tiny radius disables the bound filter for its pairs: rpmax = +inf
        p.clist[blk * P24_THREADS + rank] = make_float4(pcx, pcy, rpmin < 0.25f ? INFINITY : rpmax, __int_as_float(a));
    }
    if (tid == 0) p.ccount[blk] = total;
    if (active) p.claim_cnt[(long long)b * p.A + a] = 0;

    objpart = warp_sum_d(objpart);
    if (lane == 0) s_red[warp] = objpart;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < P24_WARPS; ++w) t += s_red[w];
        p.obj_part[blk] = t;
    }
}

// -------------------------------------------------------------------------------------------
// k_gt_match
// -------------------------------------------------------------------------------------------
#define HIT_CAP 4096
#define SURV_CAP (HIT_CAP + 1024)
#define N_SEED 16

// Upper bound of the pair value as a function of t = rpmax + d: any ray has loss <= max(1, 2 - 4 rg^2 / (rg + rp + d)^2)
// (nested rays: loss <= 1; partial and apart rays: loss <= 2 - uni/cs; DESIGN.md "top-10 bracket").
// Evaluated by lanes 0..23 of one warp.
__device__ __forceinline__ float bound_H(float rg_lane, float t, int lane) {
    float term = 0.0f;
    if (lane < P24_RAYS) {
        const float q = rg_lane + t;
        term = fmaxf(1.0f, 2.0f - __fdividef(4.0f * rg_lane * rg_lane, q * q));
    }
    return warp_sum(term) * (1.0f / 48.0f);
}

// select the `want` largest of s_surv[0..n) into s_top (descending); block-wide, returns count
__device__ int select_top(float* s_surv, int n, int want, float* s_top, KV* s_kv) {
    const int m = min(want, n);
    for (int r = 0; r < m; ++r) {
        KV best = {P24_NEG_INF, 0x7fffffff};
        for (int i = threadIdx.x; i < n; i += P24_THREADS) {
            const float v = s_surv[i];
            if (kv_gt(v, i, best.v, best.i)) {
                best.v = v;
                best.i = i;
            }
        }
        best = block_select<true>(best, s_kv);
        if (threadIdx.x == 0) {
            s_top[r] = best.v;
            if (best.i < n) s_surv[best.i] = P24_NEG_INF;
        }
        __syncthreads();
    }
    return m;
}

struct MatchShared {
    float rec[GT_REC];
    int hit[HIT_CAP];
    float surv[SURV_CAP];
    float top[P24_TOPK];
    int seed[N_SEED];
    float seedv[N_SEED];
    KV kv[P24_WARPS];
    float wmax[P24_WARPS];
    int cnt, nhit, nsurv, k, slow;
    float T, tau;
    int wanchor[P24_VCAP];
    float wcost[P24_VCAP];
    float wval[P24_VCAP];
};

// Exact top-kc sum when the bracket is not conclusive: evaluate every pair the bound cannot exclude.
__device__ __noinline__ float exact_topk_sum(const Params& p, MatchShared& S, int b, int kc, float T_seed) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float gcx = S.rec[GT_CX], gcy = S.rec[GT_CY];
    const float* img = p.outputs + (long long)b * p.img_stride;
    if (warp == 0) {
        float tau = P24_NEG_INF;
        const bool filter = !(p.flags & P24_F_NO_FILTER) && T_seed > P24_NEG_INF && S.rec[GT_RGMIN] >= 0.25f;
        if (filter) {
            const float target = T_seed - 2e-5f;
            const float rgl = lane < P24_RAYS ? S.rec[GT_RG + lane] : 0.0f;
            float lo = 0.0f, hi = 65536.0f;
            if (bound_H(rgl, lo, lane) <= target) {
                for (int it = 0; it < 26; ++it) {
                    const float mid = 0.5f * (lo + hi);
                    if (bound_H(rgl, mid, lane) <= target) lo = mid;
                    else hi = mid;
                }
                tau = lo - 0.01f - 1e-4f * lo;
            }
        }
        if (lane == 0) {
            S.T = filter ? T_seed : P24_NEG_INF;
            S.tau = tau;
            S.nsurv = 0;
        }
    }
    __syncthreads();
    const float T = S.T, tau = S.tau;
    for (int t0 = 0; t0 < p.tiles; t0 += HIT_CAP / P24_THREADS) {
        if (tid == 0) S.nhit = 0;
        __syncthreads();
        const int t1 = min(t0 + HIT_CAP / P24_THREADS, p.tiles);
        for (int tl = t0 + warp; tl < t1; tl += P24_WARPS) {
            const long long blk = (long long)b * p.tiles + tl;
            const int c = p.ccount[blk];
            for (int i = lane; i < c; i += 32) {
                const float4 q = p.clist[blk * P24_THREADS + i];
                const float dx = gcx - q.x, dy = gcy - q.y;
                const float t = q.z + sqrtf(fmaf(dx, dx, dy * dy));
                if (t >= tau || !(t == t)) S.hit[atomicAdd(&S.nhit, 1)] = __float_as_int(q.w);
            }
        }
        __syncthreads();
        const int nhit = S.nhit;
        for (int i = tid; i < nhit; i += P24_THREADS) {
            const float v = pair_value_row(S.rec, img + (long long)S.hit[i] * p.row_stride);
            if (v >= T || !(v == v)) S.surv[atomicAdd(&S.nsurv, 1)] = v;
        }
        __syncthreads();
        if (S.nsurv > 1024 && t1 < p.tiles) {
            const int m = select_top(S.surv, S.nsurv, kc, S.top, S.kv);
            if (tid < m) S.surv[tid] = S.top[tid];
            if (tid == 0) S.nsurv = m;
            __syncthreads();
        }
    }
    const int m = select_top(S.surv, S.nsurv, kc, S.top, S.kv);
    float ksum = 0.0f;
    for (int i = 0; i < m; ++i) ksum = ksum + S.top[i];
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
        const int cc = p.ccount[blk];
        for (int i = tid; i < cc; i += P24_THREADS) {
            const int a = __float_as_int(p.clist[blk * P24_THREADS + i].w);
            bool isvalid = false;
            for (int j = 0; j < nwin; ++j) isvalid |= (S.wanchor[j] == a && !(S.wval[j] < 0.0f));
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
        const KV win = block_select<false>(head, S.kv);
        if (win.i == 0x7fffffff) break;  // fewer candidates than needed
        if (li[0] == win.i && lv[0] == win.v) {
            const long long o = (long long)b * p.A + win.i;
            atomicAdd(&p.claim_cnt[o], 1);
            p.claim_gt[o] = g;
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

__global__ void __launch_bounds__(P24_THREADS) k_gt_match(Params p) {
    const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int n = p.num_gt[b];
    if (g >= n) {
        if (tid == 0) p.dyn_k[b * p.Lmax + g] = 0;
        return;
    }
    __shared__ MatchShared S;

    if (tid < GT_REC) S.rec[tid] = p.gt_rec[((long long)b * p.Lmax + g) * GT_REC + tid];
    if (tid == 0) S.cnt = 0;
    __syncthreads();
    const float gcx = S.rec[GT_CX], gcy = S.rec[GT_CY];
    const float* img = p.outputs + (long long)b * p.img_stride;

    // ---- scan the image's candidates: count, largest t, two seeds per warp -----------------------------
    float t1 = P24_NEG_INF, t2 = P24_NEG_INF;
    int a1 = 0x7fffffff, a2 = 0x7fffffff, cnt = 0;
    for (int tl = warp; tl < p.tiles; tl += P24_WARPS) {
        const long long blk = (long long)b * p.tiles + tl;
        const int c = p.ccount[blk];
        cnt += c;  // every lane of the warp holds the same running count
        for (int i = lane; i < c; i += 32) {
            const float4 q = p.clist[blk * P24_THREADS + i];
            const float dx = gcx - q.x, dy = gcy - q.y;
            const float t = q.z + sqrtf(fmaf(dx, dx, dy * dy));
            const int a = __float_as_int(q.w);
            if (kv_gt(t, a, t1, a1)) {
                t2 = t1;
                a2 = a1;
                t1 = t;
                a1 = a;
            } else if (kv_gt(t, a, t2, a2)) {
                t2 = t;
                a2 = a;
            }
        }
    }
    {
        if (lane == 0 && cnt) atomicAdd(&S.cnt, cnt);
        const KV w1 = warp_select<true>(KV{t1, a1});
        const bool owner = (a1 == w1.i) && (a1 != 0x7fffffff);
        const KV w2 = warp_select<true>(owner ? KV{t2, a2} : KV{t1, a1});
        if (lane == 0) {
            S.seed[2 * warp] = (w1.i != 0x7fffffff) ? w1.i : -1;
            S.seed[2 * warp + 1] = (w2.i != 0x7fffffff) ? w2.i : -1;
            S.wmax[warp] = w1.v;
        }
    }
    __syncthreads();
    const int ncand = S.cnt;
    const int kc = min(P24_TOPK, ncand);  // losses.py:452
    if (tid < N_SEED) {
        const int sa = S.seed[tid];
        S.seedv[tid] = (sa >= 0) ? pair_value_row(S.rec, img + (long long)sa * p.row_stride) : P24_NEG_INF;
    }
    __syncthreads();
    // ---- bracket the top-10 sum: L = sum of the 10 best seed values <= S <= 10 * H(t_max) = U ------------
    if (warp == 0) {
        float T = P24_NEG_INF, L = 0.0f;
        if (lane == 0) {
            float sv[N_SEED];
            int ns = 0;
            for (int i = 0; i < N_SEED; ++i) {
                const float v = S.seedv[i];
                if (v > P24_NEG_INF) {
                    int j = ns++;
                    while (j > 0 && sv[j - 1] < v) {
                        sv[j] = sv[j - 1];
                        --j;
                    }
                    sv[j] = v;
                }
            }
            if (kc == P24_TOPK && ns >= P24_TOPK) {
                T = sv[P24_TOPK - 1];
                for (int i = 0; i < P24_TOPK; ++i) L = L + sv[i];
            }
        }
        T = __shfl_sync(0xffffffffu, T, 0);
        L = __shfl_sync(0xffffffffu, L, 0);
        float tmax = lane < P24_WARPS ? S.wmax[lane] : P24_NEG_INF;
        tmax = warp_max(tmax);
        int slow = 1, k = 0;
        if (T > P24_NEG_INF && !(p.flags & P24_F_NO_FILTER) && S.rec[GT_RGMIN] >= 0.25f && tmax < 60000.0f) {
            const float rgl = lane < P24_RAYS ? S.rec[GT_RG + lane] : 0.0f;
            const float U = 10.0f * (bound_H(rgl, tmax * 1.0001f + 0.01f, lane) + 2e-5f);
            const float fl = floorf(L - 1e-4f), fu = floorf(U + 1e-4f);
            if (fl == fu && fl >= 1.0f) {
                slow = 0;
                k = (int)fl;
            }
        }
        if (lane == 0) {
            S.slow = slow;
            S.k = k;
            S.T = T;
        }
    }
    __syncthreads();
    int k;
    if (S.slow) {
        const float ksum = exact_topk_sum(p, S, b, kc, S.T);
        k = (int)ksum;  // dynamic k = clamp(int(sum of the top-kc values), min=1)   losses.py:454-456
        if (k < 1) k = 1;
    } else {
        k = S.k;
    }
    k = min(k, ncand);  // torch.topk would raise beyond the candidate count; clamp instead
    if (tid == 0) p.dyn_k[b * p.Lmax + g] = k;

    // ---- the GT's centre-window anchors: polygon test, exact pair value and cost of the valid ones --------
    const int wslot = b * p.Lmax + g;
    const int nwin = min(p.wcount[wslot], P24_VCAP);
    __syncthreads();
    if (tid == 0) p.wcount[wslot] = 0;  // leave the list empty for the next call
    if (tid < P24_VCAP) {
        int a = 0x7fffffff;
        float v = -1.0f;  // -1: not valid
        if (tid < nwin) {
            a = p.wlist[(long long)wslot * P24_VCAP + tid];
            const float st = p.strides[a];
            const float xc = p24_anchor_centre(p.x_shifts[a], st);
            const float yc = p24_anchor_centre(p.y_shifts[a], st);
            const bool in = (p.flags & P24_F_NO_PRUNE) ? (p24_angle_sum(S.rec + GT_VX, S.rec + GT_VY, xc, yc) >= 350.0f)
                                                       : p24_in_polygon(S.rec + GT_VX, S.rec + GT_VY, xc, yc);
            if (in) v = pair_value_row(S.rec, img + (long long)a * p.row_stride);
        }
        S.wanchor[tid] = a;
        S.wval[tid] = v;
        S.wcost[tid] = P24_POS_INF;
    }
    __syncthreads();
    {
        const int c = gt_class(S.rec, p.nc);
        for (int i = warp; i < nwin; i += P24_WARPS) {
            const float v = S.wval[i];
            if (v < 0.0f) continue;  // not in the polygon
            const float* row = img + (long long)S.wanchor[i] * p.row_stride;
            const float eo1 = 1.0f + expf(-row[26]);
            const float neg = warp_cls_neg_sum(row + 27, p.nc, eo1);
            if (lane == 0) {
                float cost = p24_cost(cls_cost_from(neg, row[27 + c], 1.0f / eo1), v, true);
                if (!(cost == cost)) cost = 3.0e38f;  // NaN inputs: keep the pair selectable, last
                S.wcost[i] = cost;
            }
        }
    }
    __syncthreads();
    int nvalid = 0;
    if (warp == 0) {
        for (int i = lane; i < nwin; i += 32) nvalid += (S.wcost[i] < P24_POS_INF) ? 1 : 0;
        nvalid = warp_sum_i(nvalid);
        const int take = min(k, nvalid);
        for (int r = 0; r < take; ++r) {
            KV best = {P24_POS_INF, 0x7fffffff};
            int bslot = -1;
            for (int i = lane; i < nwin; i += 32) {
                if (kv_lt(S.wcost[i], S.wanchor[i], best.v, best.i)) {
                    best.v = S.wcost[i];
                    best.i = S.wanchor[i];
                    bslot = i;
                }
            }
            const KV win = warp_select<false>(best);
            if (win.i == 0x7fffffff) break;
            if (bslot >= 0 && best.i == win.i && best.v == win.v) {
                const long long o = (long long)b * p.A + win.i;
                atomicAdd(&p.claim_cnt[o], 1);
                p.claim_gt[o] = g;
                S.wcost[bslot] = P24_POS_INF;
            }
            __syncwarp();
        }
        if (lane == 0) S.cnt = nvalid;
    }
    __syncthreads();
    nvalid = S.cnt;
    if (k > nvalid) spill_claims(p, S, b, g, nwin, k - nvalid);
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

// Anchor claimed by several GTs: argmin of the cost over ALL GTs (losses.py:471-476); first index on ties.
// One warp per anchor, lanes over GTs (valid pairs always beat penalised ones).
__device__ int resolve_conflict(const Params& p, const float* recs, int n, const float* row, int a) {
    const int lane = threadIdx.x & 31;
    const float st = p.strides[a];
    const float xc = p24_anchor_centre(p.x_shifts[a], st);
    const float yc = p24_anchor_centre(p.y_shifts[a], st);
    const float eo1 = 1.0f + expf(-row[26]);
    const float neg = warp_cls_neg_sum(row + 27, p.nc, eo1);
    const float obj_sig = 1.0f / eo1;
    KV best = {P24_POS_INF, 0x7fffffff};
    for (int g0 = 0; g0 < n; g0 += 32) {
        const int g = g0 + lane;
        if (g < n) {
            const float* rec = recs + g * GT_REC;
            bool valid = p24_in_centre(rec[GT_CX], rec[GT_CY], xc, yc, st);
            if (valid)
                valid = (p.flags & P24_F_NO_PRUNE) ? (p24_angle_sum(rec + GT_VX, rec + GT_VY, xc, yc) >= 350.0f)
                                                   : p24_in_polygon(rec + GT_VX, rec + GT_VY, xc, yc);
            if (valid) {
                const float v = pair_value_row(rec, row);
                float c = p24_cost(cls_cost_from(neg, row[27 + gt_class(rec, p.nc)], obj_sig), v, true);
                if (!(c == c)) c = 3.0e38f;
                if (kv_lt(c, g, best.v, best.i)) {
                    best.v = c;
                    best.i = g;
                }
            }
        }
    }
    best = warp_select<false>(best);
    if (best.i != 0x7fffffff) return best.i;
    // no valid pair at all (every claim came from a spill): argmin over the penalised costs
    for (int g0 = 0; g0 < n; g0 += 32) {
        const int g = g0 + lane;
        if (g < n) {
            const float* rec = recs + g * GT_REC;
            const float v = pair_value_row(rec, row);
            const float c = p24_cost(cls_cost_from(neg, row[27 + gt_class(rec, p.nc)], obj_sig), v, false);
            if (kv_lt(c, g, best.v, best.i)) {
                best.v = c;
                best.i = g;
            }
        }
    }
    best = warp_select<false>(best);
    return best.i != 0x7fffffff ? best.i : 0;
}

__global__ void __launch_bounds__(P24_THREADS) k_resolve_loss(Params p) {
    const int b = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int a = tile * P24_THREADS + tid;
    const int n = p.num_gt[b];
    __shared__ int s_fg[P24_THREADS];    // a_local | g << 8
    __shared__ int s_conf[P24_THREADS];  // a_local of anchors claimed by several GTs
    __shared__ int s_nfg, s_nconf;
    __shared__ double s_acc[P24_WARPS][28];
    __shared__ float s_sums[28];
    __shared__ bool s_last;
    if (tid == 0) {
        s_nfg = 0;
        s_nconf = 0;
    }
    __syncthreads();

    const float* img = p.outputs + (long long)b * p.img_stride;
    const float* recs = p.gt_rec + (long long)b * p.Lmax * GT_REC;
    if (a < p.A) {
        const long long o = (long long)b * p.A + a;
        const int cnt = n > 0 ? p.claim_cnt[o] : 0;
        if (cnt == 1) {
            const int g = p.claim_gt[o];
            p.fg_mask[o] = 1;
            p.matched_gt[o] = g;
            s_fg[atomicAdd(&s_nfg, 1)] = tid | (g << 8);
        } else if (cnt > 1) {
            s_conf[atomicAdd(&s_nconf, 1)] = tid;
        } else {
            p.fg_mask[o] = 0;
            p.matched_gt[o] = -1;
            p.pred_iou[o] = 0.0f;
        }
    }
    __syncthreads();
    const int nconf = s_nconf;
    for (int i = warp; i < nconf; i += P24_WARPS) {
        const int al = s_conf[i];
        const int aa = tile * P24_THREADS + al;
        const int g = resolve_conflict(p, recs, n, img + (long long)aa * p.row_stride, aa);
        if (lane == 0) {
            const long long o = (long long)b * p.A + aa;
            p.fg_mask[o] = 1;
            p.matched_gt[o] = g;
            s_fg[atomicAdd(&s_nfg, 1)] = al | (g << 8);
        }
    }
    __syncthreads();
    const int nfg = s_nfg;
    if (tid == 0 && nfg) atomicAdd(&p.num_fg[b], nfg);

    // ---- loss terms of the foreground anchors: one warp per anchor, lanes over rays / classes --------
    double acc_ray = 0.0;  // lane k < 24: sum of loss24[:, k]
    double acc_obj = 0.0;  // lane 0: -sum of obj logits at fg
    double acc_cls = 0.0;  // lane 0: cls BCE
    for (int i = warp; i < nfg; i += P24_WARPS) {
        const int al = s_fg[i] & 0xFF, g = s_fg[i] >> 8;
        const int aa = tile * P24_THREADS + al;
        const float* rec = recs + g * GT_REC;
        const float* row = img + (long long)aa * p.row_stride;
        const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], row[0], row[1]);
        float l = 0.0f;
        if (lane < P24_RAYS) l = p24_ray_loss(rec[GT_RG + lane], row[2 + lane], d);
        acc_ray += (double)l;
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < P24_RAYS; ++k) s = s + __shfl_sync(0xffffffffu, l, k);
        const float v = (s / 24.0f) / 2.0f;  // pair value == pred_ious_this_matching (losses.py:491)
        if (lane == 0) {
            p.pred_iou[(long long)b * p.A + aa] = v;
            acc_obj -= (double)row[26];
        }
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
            if (lane == 0) acc_cls += ((double)logf(prod) + (double)big) - (double)row[27 + c] * (double)v;
        }
    }
    if (!p.sums28) return;
    if (lane < P24_RAYS) s_acc[warp][lane] = acc_ray;
    if (lane == 0) {
        s_acc[warp][24] = acc_obj;
        s_acc[warp][25] = acc_cls;
    }
    __syncthreads();
    const int blk = b * p.tiles + tile;
    if (tid < 26) {
        double t = 0.0;
        for (int w = 0; w < P24_WARPS; ++w) t += s_acc[w][tid];
        if (tid == 24) t += p.obj_part[blk];
        p.loss_part[(long long)blk * 28 + tid] = t;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned done = atomicAdd(p.ticket, 1u);
        s_last = (done == (unsigned)(gridDim.x * gridDim.y) - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: fixed-order reduction of the partials -> deterministic sums
    const int nblk = gridDim.x * gridDim.y;
    if (warp < 7) {
        for (int q = 0; q < 4; ++q) {
            const int col = warp * 4 + q;
            if (col < 26) {
                double t = 0.0;
                for (int i = lane; i < nblk; i += 32) t += ((volatile double*)p.loss_part)[(long long)i * 28 + col];
                t = warp_sum_d(t);
                if (lane == 0) s_sums[col] = (float)t;
            } else {
                int t = 0;
                const volatile int32_t* src = (col == 26) ? p.num_fg : p.num_gt;
                for (int i = lane; i < p.B; i += 32) t += src[i];
                t = warp_sum_i(t);
                if (lane == 0) s_sums[col] = (float)t;
            }
        }
    }
    __syncthreads();
    if (tid < 28) p.sums28[tid] = s_sums[tid];
    if (tid == 0) *p.ticket = 0u;  // ready for the next call
    if (p.state26 && warp == 0) finalize_warp(s_sums, p.state26, p.result54, p.weights27);
}

__global__ void k_finalize(const float* __restrict__ sums28, float* __restrict__ state26, float* __restrict__ result54,
                           float* __restrict__ weights_n27) {
    finalize_warp(sums28, state26, result54, weights_n27);
}

size_t anchor_pass_smem(int Lmax) { return (size_t)Lmax * GT_REC * sizeof(float); }

// optional per-stage timing (profiling aid for bench.py; process-global, not thread-safe)
#define N_STAGES 3
bool g_prof_on = false;
cudaEvent_t g_prof_ev[N_STAGES + 1];
bool g_prof_have = false;
inline void prof_mark(int i, cudaStream_t st) {
    if (g_prof_on) cudaEventRecord(g_prof_ev[i], st);
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
                                     const float* strides, uint8_t* fg_mask, int32_t* matched_gt, float* pred_iou,
                                     int32_t* num_fg, int32_t* num_gt, int32_t* dyn_k, float* sums28, float* state26,
                                     float* result54, float* weights_n27, void* workspace, size_t workspace_bytes,
                                     uint32_t flags, void* stream) {
    if (!outputs || !labels || !x_shifts || !y_shifts || !strides || !fg_mask || !matched_gt || !pred_iou || !num_fg ||
        !num_gt || !dyn_k || !workspace)
        return P24_E_BADARG;
    if (B <= 0 || A <= 0 || Lmax <= 0 || num_classes <= 0 || Lmax > 65535 || B > 65535) return P24_E_BADARG;
    if (state26 && (!sums28 || !result54 || !weights_n27)) return P24_E_BADARG;
    const P24Workspace L = p24_layout(B, A, Lmax);
    if (workspace_bytes < L.total) return P24_E_WORKSPACE;
    if (((uintptr_t)workspace & 255) != 0) return P24_E_BADARG;
    const size_t dyn = anchor_pass_smem(Lmax);
    if (dyn > 160 * 1024) return P24_E_UNSUPPORTED;
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
    p.ccount = (int*)(ws + L.ccount);
    p.wcount = (int*)(ws + L.wcount);
    p.wlist = (int*)(ws + L.wlist);
    p.claim_cnt = (int*)(ws + L.claim_cnt);
    p.claim_gt = (int*)(ws + L.claim_gt);
    p.obj_part = (double*)(ws + L.obj_part);
    p.loss_part = (double*)(ws + L.loss_part);
    p.ticket = (unsigned*)(ws + L.ticket);
    p.err_flag = (int*)(ws + L.err_flag);
    p.flags = flags;
    p.tiles = p24_tiles(A);
    cudaStream_t st = (cudaStream_t)stream;

    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k_anchor_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_done = true;
    }
    prof_mark(0, st);
    k_anchor_pass<<<dim3(p.tiles, B), P24_THREADS, dyn, st>>>(p);
    prof_mark(1, st);
    k_gt_match<<<dim3(Lmax, B), P24_THREADS, 0, st>>>(p);
    prof_mark(2, st);
    k_resolve_loss<<<dim3(p.tiles, B), P24_THREADS, 0, st>>>(p);
    prof_mark(3, st);
    return (int)cudaGetLastError();
}

extern "C" int p24_loss_finalize(const float* sums28, float* state26, float* result54, float* weights_n27,
                                 void* stream) {
    if (!sums28 || !state26 || !result54 || !weights_n27) return P24_E_BADARG;
    k_finalize<<<1, 32, 0, (cudaStream_t)stream>>>(sums28, state26, result54, weights_n27);
    return (int)cudaGetLastError();
}

extern "C" int p24_profile_enable(int on) {
    if (on && !g_prof_have) {
        for (int i = 0; i <= N_STAGES; ++i) {
            const cudaError_t e = cudaEventCreate(&g_prof_ev[i]);
            if (e != cudaSuccess) return (int)e;
        }
        g_prof_have = true;
    }
    g_prof_on = on != 0;
    return 0;
}

extern "C" int p24_profile_read(float* h_ms3) {
    if (!g_prof_have || !h_ms3) return P24_E_BADARG;
    cudaError_t e = cudaEventSynchronize(g_prof_ev[N_STAGES]);
    if (e != cudaSuccess) return (int)e;
    for (int i = 0; i < N_STAGES; ++i) {
        e = cudaEventElapsedTime(&h_ms3[i], g_prof_ev[i], g_prof_ev[i + 1]);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}
