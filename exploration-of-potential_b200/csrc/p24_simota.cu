// p24_simota.cu — the fused YOLOX-24p SimOTA assignment + loss-sum path for sm_100a.
//
// Replaces, for a whole batch and without ever writing the G x A cost matrix to memory:
//   Loss_Function.get_assignments / get_in_boxes_info / pts_in_poly   models/losses.py:359-592
//   utils.boxes.bboxes_iou + pairwise circle_inter                   utils/boxes.py:102-243
//   Loss_Function.dynamic_k_matching                                 models/losses.py:444-494
//   the loss sums of Loss_Function.forward                           models/losses.py:246-302
//
// Kernel chain (all on the caller's stream, no host synchronisation):
//   k_gt_prep      one CTA per image: nlabel, per-GT records (vertices, ray lengths, safe accept /
//                  reject radii for the polygon test), zeroes the per-GT lists
//   k_anchor_pass  one CTA per 256-anchor tile: the ONE pass over the head output.  Candidate mask
//                  (polygon test OR centre window), per-anchor radius range, the valid
//                  (in polygon AND in window) pairs with their exact pair value and cost, the
//                  per-anchor argmin over valid pairs, and the sum of BCEWithLogits(obj, 0)
//   k_gt_match     one CTA per GT: exact top-10-largest pair values over the candidates (pruned with
//                  a monotone upper bound so only a few hundred pairs are evaluated) -> dynamic k;
//                  dynamic-k smallest costs from the GT's valid list (spill into the penalised
//                  regime when the list is too short) -> claims
//   k_resolve_loss one CTA per tile: conflict resolution, fg_mask / matched_gt / pred_iou, and the
//                  28 loss sums (24 per-ray GIoU losses, obj BCE, cls BCE, num_fg, num_gt)
//
// Compile with -fmad=false: the discrete decisions hang on fp32 thresholds evaluated in the
// reference's operation order; the pruning bounds use explicit fmaf.
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
    // workspace
    float* gt_rec;
    float4* anc4;
    int* vcount;
    int* vanchor;
    float* vcost;
    unsigned long long* best_key;
    int* claim_cnt;
    int* claim_gt;
    double* obj_part;
    double* loss_part;
    unsigned* ticket;
    int* err_flag;
    unsigned flags;
    int tiles;
};

#define NO_KEY 0xFFFFFFFFFFFFFFFFull

// -------------------------------------------------------------------------------------------
// k_gt_prep
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gt_prep(Params p) {
    const int b = blockIdx.x;
    const float* lab = p.labels + (long long)b * p.lab_img_stride;
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // nlabel = (labels.sum(2) > 0).sum(1)   losses.py:190
    int local = 0;
    for (int r = threadIdx.x; r < p.Lmax; r += blockDim.x) {
        const float* row = lab + (long long)r * p.lab_row_stride;
        double s = 0.0;
        for (int c = 0; c < 51; ++c) s += (double)row[c];
        if ((float)s > 0.0f) ++local;
    }
    if (local) atomicAdd(&s_n, local);
    __syncthreads();
    const int n = (p.flags & P24_F_ALL_ROWS) ? p.Lmax : s_n;
    if (threadIdx.x == 0) {
        p.num_gt[b] = n;
        p.num_fg[b] = 0;
        if (b == 0) {
            *p.ticket = 0u;
            *p.err_flag = 0;
        }
    }
    for (int g = threadIdx.x; g < p.Lmax; g += blockDim.x) {
        p.vcount[b * p.Lmax + g] = 0;
        p.dyn_k[b * p.Lmax + g] = 0;
    }
    // the first n rows are the GTs (losses.py:219-220), whatever their content
    for (int g = threadIdx.x; g < n; g += blockDim.x) {
        const float* row = lab + (long long)g * p.lab_row_stride;
        float* rec = p.gt_rec + ((long long)b * p.Lmax + g) * GT_REC;
        const float cx = row[1], cy = row[2];
        rec[GT_CX] = cx;
        rec[GT_CY] = cy;
        rec[GT_CLS] = row[0];
        float rgmax = 0.0f, rgmin = INFINITY;
        double perim = 0.0, rin = 1e30;
        bool inside = false;
        for (int k = 0; k < P24_RAYS; ++k) {
            const float x = row[3 + 2 * k], y = row[4 + 2 * k];
            const int k2 = (k + 1) % P24_RAYS;
            const float x2 = row[3 + 2 * k2], y2 = row[4 + 2 * k2];
            rec[GT_VX + k] = x;
            rec[GT_VY + k] = y;
            const float rg = p24_gt_radius(x - cx, y - cy);
            rec[GT_RG + k] = rg;
            rgmax = fmaxf(rgmax, rg);
            rgmin = fminf(rgmin, rg);
            const double ex = (double)x2 - x, ey = (double)y2 - y;
            const double len2 = ex * ex + ey * ey;
            perim += sqrt(len2);
            // distance from the centre to the edge segment
            const double wx = (double)cx - x, wy = (double)cy - y;
            double tt = len2 > 0.0 ? (wx * ex + wy * ey) / len2 : 0.0;
            tt = fmin(fmax(tt, 0.0), 1.0);
            const double qx = wx - tt * ex, qy = wy - tt * ey;
            rin = fmin(rin, sqrt(qx * qx + qy * qy));
            // crossing-number parity of the centre
            if ((y > cy) != (y2 > cy)) {
                const double xi = ((double)x2 - x) * ((double)cy - y) / ((double)y2 - y) + x;
                if ((double)cx < xi) inside = !inside;
            }
        }
        // A point inside a closed polygon has |winding| >= 1, so its total unsigned angle is
        // >= 360 degrees: the disc of radius rin around an interior centre passes the >= 350 test.
        const double ra = (inside && rin == rin) ? 0.998 * rin : 0.0;
        // Outside, the angle sum is <= perimeter / distance-to-polygon (radians):
        // it is < 349 degrees beyond rgmax + perimeter * (180/pi) / 349.
        const double rr = ((double)rgmax + perim * (57.29577951308232 / 349.0)) * 1.001 + 1e-3;
        rec[GT_RIN2] = (float)(ra * ra);
        float rrej2 = (float)(rr * rr * 1.0001);
        if (!(rrej2 == rrej2)) rrej2 = INFINITY;  // NaN labels: never reject
        rec[GT_RREJ2] = rrej2;
        rec[GT_RGMAX] = rgmax;
        rec[GT_RGMIN] = rgmin;
        rec[7] = 0.0f;
    }
}

// -------------------------------------------------------------------------------------------
// shared device helpers
// -------------------------------------------------------------------------------------------
// sum over all classes of BCE(p_j, 0), p_j = sqrt(sigmoid(cls_j) * sigmoid(obj))   losses.py:406-416
__device__ double cls_neg_sum(const float* __restrict__ cls, int nc, float obj_sig) {
    double s = 0.0;
    for (int j = 0; j < nc; ++j) s += (double)p24_bce_neg(p24_joint_prob(cls[j], obj_sig));
    return s;
}

// class cost of one (GT class, anchor) pair from the anchor's all-negative sum
__device__ __forceinline__ float cls_cost_from(double neg_sum, const float* __restrict__ cls, int c, float obj_sig) {
    const float pc = p24_joint_prob(cls[c], obj_sig);
    return (float)(neg_sum - (double)p24_bce_neg(pc) + (double)p24_bce_pos(pc));
}

// exact pair value of (GT record, prediction row in global memory)
__device__ float pair_value_row(const float* __restrict__ rec, const float* __restrict__ row) {
    const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], row[0], row[1]);
    float s = 0.0f;
#pragma unroll 4
    for (int k = 0; k < P24_RAYS; ++k) s = s + p24_ray_loss(rec[GT_RG + k], row[2 + k], d);
    return (s / 24.0f) / 2.0f;
}

// penalised cost of an arbitrary pair (slow paths only): losses.py:420-424 with ~valid
__device__ float penalised_cost(const float* __restrict__ rec, const float* __restrict__ row, int nc, double neg_sum,
                                float obj_sig) {
    const float v = pair_value_row(rec, row);
    int c = (int)rec[GT_CLS];
    c = min(max(c, 0), nc - 1);
    const float cc = cls_cost_from(neg_sum, row + 27, c, obj_sig);
    return p24_cost(cc, v, false);
}

// -------------------------------------------------------------------------------------------
// k_anchor_pass
// -------------------------------------------------------------------------------------------
#define ITEM_CAP 2048
#define G_CHUNK 8

__global__ void __launch_bounds__(P24_THREADS) k_anchor_pass(Params p) {
    extern __shared__ float s_dyn[];
    const int b = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int a = tile * P24_THREADS + tid;
    const bool active = a < p.A;
    const int n = p.num_gt[b];

    float* s_gt = s_dyn;                                  // [n * GT_REC]
    __shared__ float s_row[27 * P24_THREADS];             // transposed head rows (ch 0..26)
    __shared__ unsigned s_items[ITEM_CAP];
    __shared__ unsigned long long s_best[P24_THREADS];
    __shared__ double s_neg[P24_THREADS];
    __shared__ int s_negok[P24_THREADS];
    __shared__ int s_cand[P24_THREADS];
    __shared__ int s_nitems;
    __shared__ double s_red[P24_WARPS];

    const float* gsrc = p.gt_rec + (long long)b * p.Lmax * GT_REC;
    for (int i = tid; i < n * GT_REC; i += P24_THREADS) s_gt[i] = gsrc[i];

    const float* row = p.outputs + (long long)b * p.img_stride + (long long)(active ? a : 0) * p.row_stride;
    float pcx = 0.f, pcy = 0.f, rpmax = 0.f, rpmin = INFINITY, obj = 0.f;
    float xc = 0.f, yc = 0.f, st = 1.f;
    if (active) {
#pragma unroll
        for (int c = 0; c < 27; ++c) {
            const float v = row[c];
            s_row[c * P24_THREADS + tid] = v;
            if (c == 0) pcx = v;
            if (c == 1) pcy = v;
            if (c >= 2 && c < 26) {
                rpmax = fmaxf(rpmax, v);
                rpmin = fminf(rpmin, v);
            }
            if (c == 26) obj = v;
        }
        st = p.strides[a];
        xc = p24_anchor_centre(p.x_shifts[a], st);
        yc = p24_anchor_centre(p.y_shifts[a], st);
    }
    s_best[tid] = NO_KEY;
    s_negok[tid] = 0;
    s_cand[tid] = 0;
    double objpart = active ? (double)p24_bce_logits(obj, 0.0f) : 0.0;
    __syncthreads();

    bool cheap = false;
    if (active) {
        for (int g = 0; g < n; ++g) {
            const float* rec = s_gt + g * GT_REC;
            const float dx = rec[GT_CX] - xc, dy = rec[GT_CY] - yc;
            const float d2 = fmaf(dx, dx, dy * dy);
            cheap |= p24_in_centre(rec[GT_CX], rec[GT_CY], xc, yc, st);
            cheap |= d2 < rec[GT_RIN2];
        }
    }
    const bool no_prune = (p.flags & P24_F_NO_PRUNE) != 0;

    for (int g0 = 0; g0 < n; g0 += G_CHUNK) {
        if (tid == 0) s_nitems = 0;
        __syncthreads();
        if (active) {
            const int g1 = min(g0 + G_CHUNK, n);
            for (int g = g0; g < g1; ++g) {
                const float* rec = s_gt + g * GT_REC;
                const bool inwin = p24_in_centre(rec[GT_CX], rec[GT_CY], xc, yc, st);
                bool need = inwin;
                if (!inwin && (!cheap || no_prune)) {
                    const float dx = rec[GT_CX] - xc, dy = rec[GT_CY] - yc;
                    const float d2 = fmaf(dx, dx, dy * dy);
                    need = no_prune || d2 <= rec[GT_RREJ2];
                }
                if (need) {
                    const int slot = atomicAdd(&s_nitems, 1);
                    s_items[slot] = (unsigned)tid | ((unsigned)g << 8) | (inwin ? 0x80000000u : 0u);
                }
            }
        }
        __syncthreads();
        const int nitems = s_nitems;
        for (int i = tid; i < nitems; i += P24_THREADS) {
            const unsigned it = s_items[i];
            const int al = it & 0xFF;
            const int g = (it >> 8) & 0xFFFF;
            const bool inwin = (it & 0x80000000u) != 0;
            const float* rec = s_gt + g * GT_REC;
            const int aa = tile * P24_THREADS + al;
            const float st2 = p.strides[aa];
            const float axc = p24_anchor_centre(p.x_shifts[aa], st2);
            const float ayc = p24_anchor_centre(p.y_shifts[aa], st2);
            const float asum = p24_angle_sum(rec + GT_VX, rec + GT_VY, axc, ayc);
            if (!(asum >= 350.0f)) continue;  // losses.py:588
            s_cand[al] = 1;
            if (!inwin) continue;
            // valid pair: in polygon AND in centre window -> exact pair value and cost
            const float d = p24_centre_dist(rec[GT_CX], rec[GT_CY], s_row[al], s_row[P24_THREADS + al]);
            float s = 0.0f;
#pragma unroll 4
            for (int k = 0; k < P24_RAYS; ++k)
                s = s + p24_ray_loss(rec[GT_RG + k], s_row[(2 + k) * P24_THREADS + al], d);
            const float v = (s / 24.0f) / 2.0f;
            const float* arow = p.outputs + (long long)b * p.img_stride + (long long)aa * p.row_stride;
            const float obj_sig = p24_sigmoid(s_row[26 * P24_THREADS + al]);
            double neg;
            if (((volatile int*)s_negok)[al]) {
                neg = ((volatile double*)s_neg)[al];
            } else {
                neg = cls_neg_sum(arow + 27, p.nc, obj_sig);
                s_neg[al] = neg;
                __threadfence_block();
                s_negok[al] = 1;
            }
            int c = (int)rec[GT_CLS];
            c = min(max(c, 0), p.nc - 1);
            const float cc = cls_cost_from(neg, arow + 27, c, obj_sig);
            const float cost = p24_cost(cc, v, true);
            const int slot = atomicAdd(&p.vcount[b * p.Lmax + g], 1);
            if (slot < P24_VCAP) {
                const long long o = ((long long)b * p.Lmax + g) * P24_VCAP + slot;
                p.vanchor[o] = aa;
                p.vcost[o] = cost;
            } else {
                atomicOr(p.err_flag, 1);
            }
            atomicMin(&s_best[al], ((unsigned long long)p24_ordered(cost) << 32) | (unsigned)g);
        }
        __syncthreads();
    }

    if (active) {
        const bool cand = (n > 0) && (cheap || s_cand[tid]);
        const long long o = (long long)b * p.A + a;
        p.anc4[o] = make_float4(pcx, pcy, cand ? rpmax : -1.0f, rpmin);
        p.best_key[o] = s_best[tid];
        p.claim_cnt[o] = 0;
    }
    objpart = warp_sum_d(objpart);
    if ((tid & 31) == 0) s_red[tid >> 5] = objpart;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < P24_WARPS; ++w) t += s_red[w];
        p.obj_part[b * p.tiles + tile] = t;
    }
}

// -------------------------------------------------------------------------------------------
// k_gt_match
// -------------------------------------------------------------------------------------------
#define HIT_CAP 4096
#define SURV_CAP (HIT_CAP + 1024)
#define N_SEED 16

// Upper bound of the pair value as a function of t = rpmax + d (any ray: loss <= max(1, 2 - 4 rg^2 / (rg + rp + d)^2)
// for nested / partial / apart rays; see DESIGN.md "top-10 filter").  Evaluated by lanes 0..23 of one warp.
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

// Spill path (rare: GT with fewer valid anchors than its dynamic k): take `need` more anchors with the
// smallest PENALISED cost among the candidates that are not valid for this GT.  Ties -> lower anchor index.
__device__ __noinline__ void spill_claims(const Params& p, int b, int g, const float* rec, const int* s_vanchor,
                                          int nv, int need, KV* s_kv) {
    float lv[P24_TOPK];
    int li[P24_TOPK];
#pragma unroll
    for (int i = 0; i < P24_TOPK; ++i) {
        lv[i] = P24_POS_INF;
        li[i] = 0x7fffffff;
    }
    const float4* anc = p.anc4 + (long long)b * p.A;
    for (int a = threadIdx.x; a < p.A; a += P24_THREADS) {
        if (anc[a].z < 0.0f) continue;
        bool isvalid = false;
        for (int i = 0; i < nv; ++i) isvalid |= (s_vanchor[i] == a);
        if (isvalid) continue;
        const float* row = p.outputs + (long long)b * p.img_stride + (long long)a * p.row_stride;
        const float obj_sig = p24_sigmoid(row[26]);
        const double neg = cls_neg_sum(row + 27, p.nc, obj_sig);
        const float c = penalised_cost(rec, row, p.nc, neg, obj_sig);
        // sorted insert (ascending by (cost, anchor))
        if (kv_lt(c, a, lv[P24_TOPK - 1], li[P24_TOPK - 1])) {
            float cv = c;
            int ci = a;
#pragma unroll
            for (int i = 0; i < P24_TOPK; ++i) {
                if (kv_lt(cv, ci, lv[i], li[i])) {
                    const float tv = lv[i];
                    const int ti = li[i];
                    lv[i] = cv;
                    li[i] = ci;
                    cv = tv;
                    ci = ti;
                }
            }
        }
    }
    for (int r = 0; r < need; ++r) {
        KV head = {lv[0], li[0]};
        const KV win = block_select<false>(head, s_kv);
        if (win.i == 0x7fffffff) break;  // fewer candidates than needed
        if (li[0] == win.i && lv[0] == win.v) {
            // this thread owns the winner: claim and pop
            const long long o = (long long)b * p.A + win.i;
            atomicAdd(&p.claim_cnt[o], 1);
            p.claim_gt[o] = g;
#pragma unroll
            for (int i = 0; i < P24_TOPK - 1; ++i) {
                lv[i] = lv[i + 1];
                li[i] = li[i + 1];
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
    if (g >= n) return;

    __shared__ float s_rec[GT_REC];
    __shared__ int s_hit[HIT_CAP];
    __shared__ float s_surv[SURV_CAP];
    __shared__ float s_top[P24_TOPK];
    __shared__ int s_seed[N_SEED];
    __shared__ float s_seedv[N_SEED];
    __shared__ KV s_kv[P24_WARPS];
    __shared__ int s_cnt, s_nhit, s_nsurv;
    __shared__ float s_T, s_tau;
    __shared__ int s_vanchor[P24_VCAP];
    __shared__ float s_vcost[P24_VCAP];

    if (tid < GT_REC) s_rec[tid] = p.gt_rec[((long long)b * p.Lmax + g) * GT_REC + tid];
    if (tid == 0) {
        s_cnt = 0;
        s_nsurv = 0;
    }
    __syncthreads();
    const float gcx = s_rec[GT_CX], gcy = s_rec[GT_CY];
    const float4* anc = p.anc4 + (long long)b * p.A;
    const float* img = p.outputs + (long long)b * p.img_stride;

    // ---- pass A: candidate count and seeds (two largest t per warp) -------------------------------
    float t1 = P24_NEG_INF, t2 = P24_NEG_INF;
    int a1 = -1, a2 = -1, cnt = 0;
    for (int a = tid; a < p.A; a += P24_THREADS) {
        const float4 q = anc[a];
        if (q.z < 0.0f) continue;
        ++cnt;
        const float dx = gcx - q.x, dy = gcy - q.y;
        const float t = q.z + sqrtf(fmaf(dx, dx, dy * dy));
        if (t > t1) {
            t2 = t1;
            a2 = a1;
            t1 = t;
            a1 = a;
        } else if (t > t2) {
            t2 = t;
            a2 = a;
        }
    }
    {
        int c = cnt;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if (lane == 0 && c) atomicAdd(&s_cnt, c);
        KV w1 = warp_select<true>(KV{t1, a1 < 0 ? 0x7fffffff : a1});
        // the owner of the warp's best offers its second best in the second round
        const bool owner = (a1 >= 0 && a1 == w1.i);
        KV w2 = warp_select<true>(owner ? KV{t2, a2 < 0 ? 0x7fffffff : a2} : KV{t1, a1 < 0 ? 0x7fffffff : a1});
        if (lane == 0) {
            s_seed[2 * warp] = (w1.v > P24_NEG_INF) ? w1.i : -1;
            s_seed[2 * warp + 1] = (w2.v > P24_NEG_INF) ? w2.i : -1;
        }
    }
    __syncthreads();
    const int ncand = s_cnt;
    const int kc = min(P24_TOPK, ncand);  // losses.py:452
    if (tid < N_SEED) {
        const int sa = s_seed[tid];
        s_seedv[tid] = (sa >= 0) ? pair_value_row(s_rec, img + (long long)sa * p.row_stride) : P24_NEG_INF;
    }
    __syncthreads();
    if (warp == 0) {
        // T = 10th largest seed value (a lower bound of the 10th largest over all candidates)
        float T = P24_NEG_INF;
        if (lane == 0) {
            float sv[N_SEED];
            int ns = 0;
            for (int i = 0; i < N_SEED; ++i) {
                const float v = s_seedv[i];
                if (v > P24_NEG_INF) {
                    int j = ns++;
                    while (j > 0 && sv[j - 1] < v) {
                        sv[j] = sv[j - 1];
                        --j;
                    }
                    sv[j] = v;
                }
            }
            if (kc == P24_TOPK && ns >= P24_TOPK) T = sv[P24_TOPK - 1];
        }
        T = __shfl_sync(0xffffffffu, T, 0);
        float tau = P24_NEG_INF;
        const bool filter = !(p.flags & P24_F_NO_FILTER) && T > P24_NEG_INF && s_rec[GT_RGMIN] >= 0.25f;
        if (filter) {
            const float target = T - 2e-5f;
            const float rgl = lane < P24_RAYS ? s_rec[GT_RG + lane] : 0.0f;
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
            s_T = filter ? T : P24_NEG_INF;
            s_tau = tau;
        }
    }
    __syncthreads();
    const float T = s_T, tau = s_tau;

    // ---- pass B: exact values of the pairs the bound cannot exclude --------------------------------
    for (int base = 0; base < p.A; base += HIT_CAP) {
        if (tid == 0) s_nhit = 0;
        __syncthreads();
        const int end = min(base + HIT_CAP, p.A);
        for (int a = base + tid; a < end; a += P24_THREADS) {
            const float4 q = anc[a];
            if (q.z < 0.0f) continue;
            const float dx = gcx - q.x, dy = gcy - q.y;
            const float t = q.z + sqrtf(fmaf(dx, dx, dy * dy));
            if (t >= tau || q.w < 0.25f || !(t == t)) s_hit[atomicAdd(&s_nhit, 1)] = a;
        }
        __syncthreads();
        const int nhit = s_nhit;
        for (int i = tid; i < nhit; i += P24_THREADS) {
            const float v = pair_value_row(s_rec, img + (long long)s_hit[i] * p.row_stride);
            if (v >= T || !(v == v)) s_surv[atomicAdd(&s_nsurv, 1)] = v;
        }
        __syncthreads();
        if (s_nsurv > 1024 && end < p.A) {
            const int m = select_top(s_surv, s_nsurv, kc, s_top, s_kv);
            if (tid < m) s_surv[tid] = s_top[tid];
            if (tid == 0) s_nsurv = m;
            __syncthreads();
        }
    }
    const int m = select_top(s_surv, s_nsurv, kc, s_top, s_kv);
    // dynamic k = clamp(int(sum of the top-kc values), min=1)   losses.py:454-456
    float ksum = 0.0f;
    for (int i = 0; i < m; ++i) ksum = ksum + s_top[i];
    int k = (int)ksum;
    if (k < 1) k = 1;
    k = min(k, ncand);  // torch.topk would raise beyond the candidate count; clamp instead
    if (tid == 0) p.dyn_k[b * p.Lmax + g] = k;

    // ---- the k smallest costs of the GT's valid list ----------------------------------------------
    const int nv = min(p.vcount[b * p.Lmax + g], P24_VCAP);
    if (tid < P24_VCAP) {
        const long long o = ((long long)b * p.Lmax + g) * P24_VCAP + tid;
        s_vanchor[tid] = tid < nv ? p.vanchor[o] : 0x7fffffff;
        s_vcost[tid] = tid < nv ? p.vcost[o] : P24_POS_INF;
    }
    __syncthreads();
    if (warp == 0) {
        const int take = min(k, nv);
        for (int r = 0; r < take; ++r) {
            KV best = {P24_POS_INF, 0x7fffffff};
            int bslot = -1;
            for (int i = lane; i < nv; i += 32) {
                if (kv_lt(s_vcost[i], s_vanchor[i], best.v, best.i)) {
                    best.v = s_vcost[i];
                    best.i = s_vanchor[i];
                    bslot = i;
                }
            }
            const KV win = warp_select<false>(best);
            if (bslot >= 0 && best.i == win.i && best.v == win.v) {
                const long long o = (long long)b * p.A + win.i;
                atomicAdd(&p.claim_cnt[o], 1);
                p.claim_gt[o] = g;
                s_vcost[bslot] = P24_POS_INF;
                s_vanchor[bslot] |= 0x40000000;  // keep the anchor id recoverable for the spill path
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (k > nv) {
        if (tid < nv) s_vanchor[tid] &= 0x3FFFFFFF;
        __syncthreads();
        spill_claims(p, b, g, s_rec, s_vanchor, nv, k - nv, s_kv);
    }
}

// -------------------------------------------------------------------------------------------
// k_resolve_loss
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(P24_THREADS) k_resolve_loss(Params p) {
    const int b = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int a = tile * P24_THREADS + tid;
    const int n = p.num_gt[b];
    __shared__ int s_fg[P24_THREADS];  // a_local | g << 8
    __shared__ int s_nfg;
    __shared__ double s_acc[P24_WARPS][28];
    __shared__ bool s_last;
    if (tid == 0) s_nfg = 0;
    __syncthreads();

    const float* img = p.outputs + (long long)b * p.img_stride;
    const float* recs = p.gt_rec + (long long)b * p.Lmax * GT_REC;
    if (a < p.A) {
        const long long o = (long long)b * p.A + a;
        int g = -1;
        if (n > 0) {
            const int cnt = p.claim_cnt[o];
            if (cnt == 1) {
                g = p.claim_gt[o];
            } else if (cnt > 1) {
                // claimed by several GTs: argmin of the cost over ALL GTs (losses.py:471-476); valid pairs
                // always beat penalised ones, and their argmin was recorded by k_anchor_pass
                const unsigned long long key = p.best_key[o];
                if (key != NO_KEY) {
                    g = (int)(key & 0xFFFFFFFFu);
                } else {
                    const float* row = img + (long long)a * p.row_stride;
                    const float obj_sig = p24_sigmoid(row[26]);
                    const double neg = cls_neg_sum(row + 27, p.nc, obj_sig);
                    float bestc = P24_POS_INF;
                    for (int gg = 0; gg < n; ++gg) {
                        const float c = penalised_cost(recs + gg * GT_REC, row, p.nc, neg, obj_sig);
                        if (c < bestc) {
                            bestc = c;
                            g = gg;
                        }
                    }
                    if (g < 0) g = 0;
                }
            }
        }
        p.fg_mask[o] = g >= 0 ? 1 : 0;
        p.matched_gt[o] = g;
        if (g < 0) p.pred_iou[o] = 0.0f;
        else s_fg[atomicAdd(&s_nfg, 1)] = tid | (g << 8);
    }
    __syncthreads();
    const int nfg = s_nfg;
    if (tid == 0 && nfg) atomicAdd(&p.num_fg[b], nfg);

    // ---- loss terms of the foreground anchors: one warp per anchor, lanes over rays / classes --------
    double acc_ray = 0.0;   // lane k < 24: sum of loss24[:, k]
    double acc_obj = 0.0;   // lane 0: -sum of obj logits at fg
    double acc_cls = 0.0;   // all lanes: partial cls BCE
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
        const float v = (s / 24.0f) / 2.0f;   // pair value == pred_ious_this_matching (losses.py:491)
        if (lane == 0) {
            p.pred_iou[(long long)b * p.A + aa] = v;
            acc_obj -= (double)row[26];
        }
        if (p.sums28) {
            int c = (int)rec[GT_CLS];
            c = min(max(c, 0), p.nc - 1);
            for (int j = lane; j < p.nc; j += 32)
                acc_cls += (double)p24_bce_logits(row[27 + j], j == c ? v : 0.0f);
        }
    }
    if (!p.sums28) return;
    acc_cls = warp_sum_d(acc_cls);
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
    if (tid < 26) {
        double t = 0.0;
        for (int i = 0; i < nblk; ++i) t += ((volatile double*)p.loss_part)[(long long)i * 28 + tid];
        p.sums28[tid] = (float)t;
    } else if (tid == 26) {
        long long t = 0;
        for (int i = 0; i < p.B; ++i) t += ((volatile int32_t*)p.num_fg)[i];
        p.sums28[26] = (float)t;
    } else if (tid == 27) {
        long long t = 0;
        for (int i = 0; i < p.B; ++i) t += p.num_gt[i];
        p.sums28[27] = (float)t;
    }
}

// -------------------------------------------------------------------------------------------
// k_finalize: normalisation + stateful re-weighting, losses.py:280-345 (one warp)
// -------------------------------------------------------------------------------------------
__global__ void k_finalize(const float* __restrict__ sums28, float* __restrict__ state26, float* __restrict__ result54,
                           float* __restrict__ weights_n27) {
    const int lane = threadIdx.x;
    const float nfg = fmaxf(sums28[26], 1.0f);
    const float ngt = fmaxf(sums28[27], 1.0f);
    float loss = 0.0f, e = 0.0f;
    if (lane < 26) {
        loss = sums28[lane] / nfg;                          // loss_iou[k], loss_obj, loss_cls
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
        result54[1 + lane] = wl;       // reg_w * loss_iou
        result54[28 + lane] = w;       // reg_w
        weights_n27[lane] = w;
    }
    if (lane == 24) {
        result54[25] = loss;           // loss_obj
        result54[52] = w;
        weights_n27[24] = w;
    }
    if (lane == 25) {
        result54[26] = loss;           // loss_cls
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

size_t anchor_pass_smem(int Lmax) { return (size_t)Lmax * GT_REC * sizeof(float); }

// optional per-stage timing (profiling aid for bench.py; process-global, not thread-safe)
#define N_STAGES 4
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

extern "C" int p24_simota_loss_batch(const float* outputs, int64_t img_stride, int64_t row_stride, int B, int A,
                                     int num_classes, const float* labels, int64_t lab_img_stride,
                                     int64_t lab_row_stride, int Lmax, const float* x_shifts, const float* y_shifts,
                                     const float* strides, uint8_t* fg_mask, int32_t* matched_gt, float* pred_iou,
                                     int32_t* num_fg, int32_t* num_gt, int32_t* dyn_k, float* sums28, void* workspace,
                                     size_t workspace_bytes, uint32_t flags, void* stream) {
    if (!outputs || !labels || !x_shifts || !y_shifts || !strides || !fg_mask || !matched_gt || !pred_iou || !num_fg ||
        !num_gt || !dyn_k || !workspace)
        return P24_E_BADARG;
    if (B <= 0 || A <= 0 || Lmax <= 0 || num_classes <= 0 || Lmax > 65535) return P24_E_BADARG;
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
    p.gt_rec = (float*)(ws + L.gt_rec);
    p.anc4 = (float4*)(ws + L.anc4);
    p.vcount = (int*)(ws + L.vcount);
    p.vanchor = (int*)(ws + L.vanchor);
    p.vcost = (float*)(ws + L.vcost);
    p.best_key = (unsigned long long*)(ws + L.best_key);
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
    k_gt_prep<<<B, 128, 0, st>>>(p);
    prof_mark(1, st);
    k_anchor_pass<<<dim3(p.tiles, B), P24_THREADS, dyn, st>>>(p);
    prof_mark(2, st);
    k_gt_match<<<dim3(Lmax, B), P24_THREADS, 0, st>>>(p);
    prof_mark(3, st);
    k_resolve_loss<<<dim3(p.tiles, B), P24_THREADS, 0, st>>>(p);
    prof_mark(4, st);
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

extern "C" int p24_profile_read(float* h_ms4) {
    if (!g_prof_have || !h_ms4) return P24_E_BADARG;
    cudaError_t e = cudaEventSynchronize(g_prof_ev[N_STAGES]);
    if (e != cudaSuccess) return (int)e;
    for (int i = 0; i < N_STAGES; ++i) {
        e = cudaEventElapsedTime(&h_ms4[i], g_prof_ev[i], g_prof_ev[i + 1]);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}
