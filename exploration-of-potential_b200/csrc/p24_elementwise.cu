// p24_elementwise.cu — the stand-alone entry points of the YOLOX-24p loss path for sm_100a: the element-wise
// concentric-circle IoU (IOUloss.circle_inter / IOUloss.forward, models/losses.py:23-157) with its backward, the
// pairwise pair value (utils.boxes.bboxes_iou, utils/boxes.py:166-243), dynamic_k_matching on materialised matrices
// (models/losses.py:444-494) and the backward of the whole loss (autograd of models/losses.py:283-341).
//
// The forward arithmetic follows the reference's operation order (compile with -fmad=false, see p24_math.cuh); the
// backward kernels use fused multiply-adds freely (gradients are compared at 1e-5 relative).
#include <string.h>

#include "p24_common.cuh"

namespace {

// -------------------------------------------------------------------------------------------
// forward kernels, one thread per (pair, ray)
// -------------------------------------------------------------------------------------------
__global__ void k_circle_inter(const float* __restrict__ gcx, const float* __restrict__ gcy, const float* __restrict__ gr,
                               long long gr_stride, const float* __restrict__ pcx, const float* __restrict__ pcy,
                               const float* __restrict__ pr, long long pr_stride, int n, float* __restrict__ res,
                               float* __restrict__ dist) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * P24_RAYS) return;
    const int r = (int)(i / P24_RAYS), k = (int)(i - (long long)r * P24_RAYS);
    const float d = p24_centre_dist(gcx[r], gcy[r], pcx[r], pcy[r]);
    float inter;
    p24_ray_loss(gr[r * gr_stride + k], pr[r * pr_stride + k], d, &inter);
    res[i] = inter;
    dist[i] = d;
}

__global__ void k_iou_loss_fwd(const float* __restrict__ pred, long long ps, const float* __restrict__ target,
                               long long ts, int n, float* __restrict__ loss24) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * P24_RAYS) return;
    const int r = (int)(i / P24_RAYS), k = (int)(i - (long long)r * P24_RAYS);
    const float* t = target + r * ts;
    const float* q = pred + r * ps;
    const float rg = p24_gt_radius(t[2 + 2 * k] - t[0], t[3 + 2 * k] - t[1]);
    const float d = p24_centre_dist(t[0], t[1], q[0], q[1]);
    loss24[i] = p24_ray_loss(rg, q[2 + k], d);
}

__global__ void k_pair_iou(const float* __restrict__ gt50, long long gs, int G, const float* __restrict__ pred,
                           long long ps, int P, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)G * P) return;
    const int g = (int)(i / P), a = (int)(i - (long long)g * P);
    const float* t = gt50 + g * gs;
    const float* q = pred + a * ps;
    const float d = p24_centre_dist(t[0], t[1], q[0], q[1]);
    float s = 0.0f;
#pragma unroll 1
    for (int k = 0; k < P24_RAYS; ++k) {
        const float rg = p24_gt_radius(t[2 + 2 * k] - t[0], t[3 + 2 * k] - t[1]);
        s = s + p24_ray_loss(rg, q[2 + k], d);
    }
    out[i] = (s / 24.0f) / 2.0f;
}

// -------------------------------------------------------------------------------------------
// d(loss24_k) / d(rp_k) and d(loss24_k) / d(d): autograd of losses.py:36-72,118-151
// (zero through the clipped acos arguments and through the branch masks, like torch)
// -------------------------------------------------------------------------------------------
__device__ void ray_loss_grad(float rg, float rp, float d, float& dl_drp, float& dl_dd) {
    const float PI = P24_PI;
    const bool pmin = rp < rg;   // torch.min / torch.max over stack((gt, pd)) return the FIRST index on ties -> gt
    const bool pmax = rp > rg;
    const float rmin = fminf(rg, rp), rmax = fmaxf(rg, rp);
    const float drmin = pmin ? 1.0f : 0.0f, drmax = pmax ? 1.0f : 0.0f;
    const bool nested = fabsf(rg - rp) >= d;
    const bool apart = d >= rg + rp;
    float inter = 0.0f, di_drp = 0.0f, di_dd = 0.0f;
    if (apart) {
        // res[mask] = 0: constant
    } else if (nested) {
        inter = PI * rmin * rmin;
        di_drp = 2.0f * PI * rmin * drmin;
    } else {
        const float n1 = rmin * rmin + d * d - rmax * rmax, D1 = 2.0f * rmin * d + 1e-8f;
        const float n2 = rmax * rmax + d * d - rmin * rmin, D2 = 2.0f * rmax * d + 1e-8f;
        const float x1 = n1 / D1, x2 = n2 / D2;
        const float dx1_drp = (2.0f * rmin * drmin - 2.0f * rmax * drmax) / D1 - n1 * (2.0f * d * drmin) / (D1 * D1);
        const float dx1_dd = (2.0f * d) / D1 - n1 * (2.0f * rmin) / (D1 * D1);
        const float dx2_drp = (2.0f * rmax * drmax - 2.0f * rmin * drmin) / D2 - n2 * (2.0f * d * drmax) / (D2 * D2);
        const float dx2_dd = (2.0f * d) / D2 - n2 * (2.0f * rmax) / (D2 * D2);
        const bool in1 = x1 >= -0.99f && x1 <= 0.99f, in2 = x2 >= -0.99f && x2 <= 0.99f;
        const float c1 = fminf(fmaxf(x1, -0.99f), 0.99f), c2 = fminf(fmaxf(x2, -0.99f), 0.99f);
        const float a1 = acosf(c1), a2 = acosf(c2);
        const float g1 = in1 ? -rsqrtf(1.0f - c1 * c1) : 0.0f, g2 = in2 ? -rsqrtf(1.0f - c2 * c2) : 0.0f;
        const float da1_drp = g1 * dx1_drp, da1_dd = g1 * dx1_dd;
        const float da2_drp = g2 * dx2_drp, da2_dd = g2 * dx2_dd;
        const float s1 = sinf(a1), co1 = cosf(a1);
        inter = a1 * rmin * rmin + a2 * rmax * rmax - rmin * d * s1;
        di_drp = da1_drp * rmin * rmin + a1 * 2.0f * rmin * drmin + da2_drp * rmax * rmax + a2 * 2.0f * rmax * drmax -
                 (drmin * d * s1 + rmin * d * co1 * da1_drp);
        di_dd = da1_dd * rmin * rmin + da2_dd * rmax * rmax - (rmin * s1 + rmin * d * co1 * da1_dd);
    }
    const float ag = PI * rg * rg, ap = PI * rp * rp;
    const float uni = ag + ap - inter;
    const float du_drp = 2.0f * PI * rp - di_drp, du_dd = -di_dd;
    const float den = uni + 1e-6f;
    const float diou_drp = (di_drp * den - inter * du_drp) / (den * den);
    const float diou_dd = (di_dd * den - inter * du_dd) / (den * den);
    const float cl = nested ? rmax : 0.5f * (rg + rp + d);
    const float dcl_drp = nested ? drmax : 0.5f, dcl_dd = nested ? 0.0f : 0.5f;
    const float cs = PI * cl * cl;
    const float dcs_drp = 2.0f * PI * cl * dcl_drp, dcs_dd = 2.0f * PI * cl * dcl_dd;
    // loss = 1 - (iou - (cs - uni) / cs) = 2 - iou - uni / cs
    dl_drp = -diou_drp - (du_drp * cs - uni * dcs_drp) / (cs * cs);
    dl_dd = -diou_dd - (du_dd * cs - uni * dcs_dd) / (cs * cs);
}

// gradient of sum_k w_k * loss24_k w.r.t. (pcx, pcy, rp[24]) of one matched pair; lanes 0..23 hold one ray each.
// Returns the lane's d/d(rp_k); gx / gy are warp-reduced (valid in every lane).
__device__ __forceinline__ float pair_grad(const float* t50, const float* q26, float w_lane, int lane, float& gx, float& gy) {
    const float dxc = t50[0] - q26[0], dyc = t50[1] - q26[1];
    const float d = sqrtf(dxc * dxc + dyc * dyc);
    float grp = 0.0f, gd = 0.0f;
    if (lane < P24_RAYS) {
        const float rg = p24_gt_radius(t50[2 + 2 * lane] - t50[0], t50[3 + 2 * lane] - t50[1]);
        float a, bb;
        ray_loss_grad(rg, q26[2 + lane], d, a, bb);
        grp = w_lane * a;
        gd = w_lane * bb;
    }
    gd = warp_sum(gd);
    // d = sqrt((gcx - pcx)^2 + (gcy - pcy)^2): dd/dpcx = -(gcx - pcx) / d (torch yields nan at d == 0; we give 0)
    const float inv = d > 0.0f ? 1.0f / d : 0.0f;
    gx = -gd * dxc * inv;
    gy = -gd * dyc * inv;
    return grp;
}

__global__ void k_iou_loss_bwd(const float* __restrict__ pred, long long ps, const float* __restrict__ target,
                               long long ts, const float* __restrict__ grad24, int n, float* __restrict__ gpred) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= n) return;
    const float w = lane < P24_RAYS ? grad24[(long long)r * P24_RAYS + lane] : 0.0f;
    float gx, gy;
    const float grp = pair_grad(target + r * ts, pred + r * ps, w, lane, gx, gy);
    float* o = gpred + (long long)r * 26;
    if (lane < P24_RAYS) o[2 + lane] = grp;
    if (lane == 0) {
        o[0] = gx;
        o[1] = gy;
    }
}

// -------------------------------------------------------------------------------------------
// backward of the whole loss w.r.t. the head output: one warp per anchor row
//   loss = sum_k w_k S_iou[k] / N + w_obj S_obj / N + w_cls S_cls / N   (weights are constants, losses.py:312-314)
// -------------------------------------------------------------------------------------------
__global__ void k_loss_bwd(const float* __restrict__ outputs, long long img_stride, long long row_stride, int B, int A,
                           int nc, const float* __restrict__ labels, long long lab_img_stride, long long lab_row_stride,
                           const uint8_t* __restrict__ fg_mask, const int32_t* __restrict__ matched_gt,
                           const float* __restrict__ pred_iou, const float* __restrict__ wn27,
                           const float* __restrict__ grad_scale, float* __restrict__ gout) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= (long long)B * A) return;
    const int b = (int)(row / A);
    const int a = (int)(row - (long long)b * A);
    const int C = 27 + nc;
    const float* x = outputs + b * img_stride + a * row_stride;
    float* g = gout + row * C;
    const float scale = (grad_scale ? grad_scale[0] : 1.0f) / wn27[26];
    const bool fg = fg_mask[row] != 0;
    // objectness: d BCEWithLogits(x, t) / dx = sigmoid(x) - t, every anchor
    if (lane == 0) g[26] = scale * wn27[24] * (p24_sigmoid(x[26]) - (fg ? 1.0f : 0.0f));
    if (!fg) {
        for (int c = lane; c < C; c += 32)
            if (c != 26) g[c] = 0.0f;
        return;
    }
    const int m = matched_gt[row];
    const float* lab = labels + b * lab_img_stride + m * lab_row_stride;
    int cls = (int)lab[0];
    cls = min(max(cls, 0), nc - 1);
    const float v = pred_iou[row];
    for (int j = lane; j < nc; j += 32) g[27 + j] = scale * wn27[25] * (p24_sigmoid(x[27 + j]) - (j == cls ? v : 0.0f));
    float gx, gy;
    const float grp = pair_grad(lab + 1, x, lane < P24_RAYS ? scale * wn27[lane] : 0.0f, lane, gx, gy);
    if (lane < P24_RAYS) g[2 + lane] = grp;
    if (lane == 0) {
        g[0] = gx;
        g[1] = gy;
    }
}

// -------------------------------------------------------------------------------------------
// the same backward w.r.t. the head's RAW per-level conv outputs (the decode of yolo_head_24p.py:233-235 folded in:
// d/d(raw centre) = stride * d/d(centre), d/d(raw radius) = radius * d/d(radius); obj / cls logits pass through).
// One CTA per 256 consecutive anchors of one image: a thread per anchor writes the planar obj / cls / zero gradients
// (coalesced across anchors), then the warps share the CTA's foreground anchors for the geometry part.
// -------------------------------------------------------------------------------------------
struct RawBwd {
    const float* in[12];   // [reg | obj | cls][level]
    long long in_bs[12];
    float* out[12];        // dense [B, C, H, W] per tensor
    int off[5], W[4];
    float st[4];
    int nlev;
};

__global__ void __launch_bounds__(256) k_loss_bwd_raw(RawBwd r, int B, int A, int nc, const float* __restrict__ labels,
                                                      long long lab_img_stride, long long lab_row_stride,
                                                      const uint8_t* __restrict__ fg_mask,
                                                      const int32_t* __restrict__ matched_gt,
                                                      const float* __restrict__ pred_iou, const float* __restrict__ wn27,
                                                      const float* __restrict__ grad_scale) {
    __shared__ int s_fg[256];
    __shared__ int s_nfg;
    __shared__ float s_q[8][26];
    const int b = blockIdx.y, a = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_nfg = 0;
    __syncthreads();
    const float scale = (grad_scale ? grad_scale[0] : 1.0f) / wn27[26];
    if (a < A) {
        int l = 0;
        for (int q = 1; q < r.nlev; ++q) l += a >= r.off[q] ? 1 : 0;
        const int i = a - r.off[l];
        const long long plane = (long long)(r.off[l + 1] - r.off[l]);
        const bool fg = fg_mask[(long long)b * A + a] != 0;
        const float xo = r.in[4 + l][b * r.in_bs[4 + l] + i];
        r.out[4 + l][b * plane + i] = scale * wn27[24] * (p24_sigmoid(xo) - (fg ? 1.0f : 0.0f));
        float* gc = r.out[8 + l] + (long long)b * nc * plane + i;
        float* gr = r.out[l] + (long long)b * 26 * plane + i;
        if (!fg) {
            for (int j = 0; j < nc; ++j) gc[j * plane] = 0.0f;
            for (int j = 0; j < 26; ++j) gr[j * plane] = 0.0f;
        } else {
            const int m = matched_gt[(long long)b * A + a];
            int cls = (int)labels[b * lab_img_stride + m * lab_row_stride];
            cls = min(max(cls, 0), nc - 1);
            const float v = pred_iou[(long long)b * A + a];
            const float* xc = r.in[8 + l] + b * r.in_bs[8 + l] + i;
            for (int j = 0; j < nc; ++j) gc[j * plane] = scale * wn27[25] * (p24_sigmoid(xc[j * plane]) - (j == cls ? v : 0.0f));
            s_fg[atomicAdd(&s_nfg, 1)] = a;
        }
    }
    __syncthreads();
    for (int f = warp; f < s_nfg; f += 8) {
        const int aa = s_fg[f];
        int l = 0;
        for (int q = 1; q < r.nlev; ++q) l += aa >= r.off[q] ? 1 : 0;
        const int i = aa - r.off[l];
        const long long plane = (long long)(r.off[l + 1] - r.off[l]);
        const float st = r.st[l];
        float dec = 0.0f;
        if (lane < 26) {
            const float x = r.in[l][b * r.in_bs[l] + lane * plane + i];
            if (lane >= 2) dec = expf(x) * st;
            else dec = (x + (float)(lane == 0 ? i % r.W[l] : i / r.W[l])) * st;
            s_q[warp][lane] = dec;
        }
        __syncwarp();
        const int m = matched_gt[(long long)b * A + aa];
        const float* lab = labels + b * lab_img_stride + m * lab_row_stride;
        float gx, gy;
        const float grp = pair_grad(lab + 1, s_q[warp], lane < P24_RAYS ? scale * wn27[lane] : 0.0f, lane, gx, gy);
        float* gr = r.out[l] + (long long)b * 26 * plane + i;
        if (lane < P24_RAYS) gr[(2 + lane) * plane] = grp * s_q[warp][2 + lane];
        if (lane == 0) {
            gr[0] = gx * st;
            gr[plane] = gy * st;
        }
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------
// label packing of the dataset's TrainTransform (datasets/data_augment.py:131-174) for a whole batch: ragged normalised
// targets [cls, cx, cy, 24 x (x, y)] in float64 (np.loadtxt) -> labels[B, max_labels, 51] fp32 in pixels of the padded
// input.  The reference computes in float64 (x * width_o, y * height_o, then * r_o) and casts once at the end: so does
// this kernel, with r_o = min(in_h / h_o, in_w / w_o) in float64 like preproc (data_augment.py:104).  One thread per
// output element; rows beyond the image's targets (and beyond max_labels) are zero.  nlabel = the count of
// losses.py:190 ((labels.sum(dim=2) > 0).sum) of the packed rows.
// -------------------------------------------------------------------------------------------
__global__ void k_pack_labels(const double* __restrict__ targets, const int32_t* __restrict__ offsets,
                              const int32_t* __restrict__ shapes, int in_h, int in_w, int B, int max_labels,
                              float* __restrict__ labels, int32_t* __restrict__ nlabel) {
    const int b = blockIdx.y;
    const int n = min(offsets[b + 1] - offsets[b], max_labels);
    const double h_o = (double)shapes[2 * b], w_o = (double)shapes[2 * b + 1];
    const double r_o = fmin((double)in_h / h_o, (double)in_w / w_o);
    float* out = labels + (long long)b * max_labels * 51;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < max_labels * 51; e += gridDim.x * blockDim.x) {
        const int row = e / 51, c = e - row * 51;
        float v = 0.0f;
        if (row < n) {
            const double t = targets[((long long)offsets[b] + row) * 51 + c];
            if (c == 0) v = (float)t;
            else v = (float)((t * (((c - 1) & 1) ? h_o : w_o)) * r_o);  // boxes[:, 0::2] *= width, [:, 1::2] *= height, then *= r
        }
        out[e] = v;
    }
    if (nlabel && blockIdx.x == 0 && threadIdx.x < 32) {
        int cnt = 0;
        for (int row = threadIdx.x; row < n; row += 32) {
            float sm = 0.0f;  // torch sums the 51 floats of a row in fp32; only the sign matters here
            for (int c = 0; c < 51; ++c) {
                const double t = targets[((long long)offsets[b] + row) * 51 + c];
                sm += c == 0 ? (float)t : (float)((t * (((c - 1) & 1) ? h_o : w_o)) * r_o);
            }
            cnt += sm > 0.0f ? 1 : 0;
        }
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (threadIdx.x == 0) nlabel[b] = cnt;
    }
}

// -------------------------------------------------------------------------------------------
// dynamic_k_matching on materialised [G, P] matrices (losses.py:444-494)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(P24_THREADS) k_dynk_select(const float* __restrict__ cost, const float* __restrict__ ious,
                                                           int G, int P, int* __restrict__ cnt, int32_t* __restrict__ matched,
                                                           int32_t* __restrict__ dyn_k, unsigned char* __restrict__ taken) {
    // one CTA per GT: top-min(10, P) largest ious -> k; the k smallest costs -> claims.  `taken` [G, P] scratch bytes.
    const int g = blockIdx.x, tid = threadIdx.x;
    __shared__ KV s_kv[P24_WARPS];
    const float* io = ious + (long long)g * P;
    const float* co = cost + (long long)g * P;
    unsigned char* tk = taken + (long long)g * P;
    for (int i = tid; i < P; i += P24_THREADS) tk[i] = 0;
    __syncthreads();
    const int kc = min(P24_TOPK, P);
    float ksum = 0.0f;
    for (int r = 0; r < kc; ++r) {
        KV best = {P24_NEG_INF, 0x7fffffff};
        for (int i = tid; i < P; i += P24_THREADS) {
            if (tk[i]) continue;
            float v = io[i];
            if (!(v == v)) v = P24_POS_INF;  // NaN sorts first in torch.topk
            if (kv_gt(v, i, best.v, best.i)) {
                best.v = v;
                best.i = i;
            }
        }
        best = block_select<true>(best, s_kv);
        if (best.i == 0x7fffffff) break;
        if (tid == 0) tk[best.i] = 1;
        ksum = ksum + (best.v == P24_POS_INF ? NAN : best.v);
        __syncthreads();
    }
    int k = (int)ksum;
    if (k < 1) k = 1;
    k = min(k, P);
    if (tid == 0) dyn_k[g] = k;
    __syncthreads();
    for (int i = tid; i < P; i += P24_THREADS) tk[i] = 0;
    __syncthreads();
    for (int r = 0; r < k; ++r) {
        KV best = {P24_POS_INF, 0x7fffffff};
        for (int i = tid; i < P; i += P24_THREADS) {
            if (tk[i]) continue;
            if (kv_lt(co[i], i, best.v, best.i)) {
                best.v = co[i];
                best.i = i;
            }
        }
        best = block_select<false>(best, s_kv);
        if (best.i == 0x7fffffff) break;
        if (tid == 0) {
            tk[best.i] = 1;
            atomicAdd(&cnt[best.i], 1);
            matched[best.i] = g;
        }
        __syncthreads();
    }
}

__global__ void k_dynk_resolve(const float* __restrict__ cost, const float* __restrict__ ious, int G, int P,
                               const int* __restrict__ cnt, uint8_t* __restrict__ fg_in, int32_t* __restrict__ matched,
                               float* __restrict__ matched_iou, int32_t* __restrict__ num_fg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int c = cnt[i];
    int g = -1;
    if (c == 1) {
        g = matched[i];
    } else if (c > 1) {  // argmin over ALL GTs, first index on ties (losses.py:474)
        float best = P24_POS_INF;
        for (int q = 0; q < G; ++q) {
            const float v = cost[(long long)q * P + i];
            if (g < 0 || v < best) {
                best = v;
                g = q;
            }
        }
    }
    fg_in[i] = g >= 0 ? 1 : 0;
    matched[i] = g;
    matched_iou[i] = g >= 0 ? ious[(long long)g * P + i] : 0.0f;
    if (g >= 0) atomicAdd(num_fg, 1);
}

__global__ void k_zero_i32(int* p, int n, int32_t* one) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
    if (i == 0 && one) *one = 0;
}

}  // namespace

// -------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------
extern "C" int p24_circle_inter_fwd(const float* gt_cx, const float* gt_cy, const float* gt_r, int64_t gt_r_stride,
                                    const float* pd_cx, const float* pd_cy, const float* pd_r, int64_t pd_r_stride, int n,
                                    float* res_inter, float* dist, void* stream) {
    if (n < 0) return P24_E_BADARG;
    if (n == 0) return 0;
    if (!gt_cx || !gt_cy || !gt_r || !pd_cx || !pd_cy || !pd_r || !res_inter || !dist) return P24_E_BADARG;
    const long long tot = (long long)n * P24_RAYS;
    k_circle_inter<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gt_cx, gt_cy, gt_r, gt_r_stride, pd_cx, pd_cy,
                                                                                 pd_r, pd_r_stride, n, res_inter, dist);
    return (int)cudaGetLastError();
}

extern "C" int p24_iou_loss_fwd(const float* pred, int64_t pred_stride, const float* target, int64_t target_stride, int n,
                                float* loss24, void* stream) {
    if (n < 0) return P24_E_BADARG;
    if (n == 0) return 0;
    if (!pred || !target || !loss24) return P24_E_BADARG;
    const long long tot = (long long)n * P24_RAYS;
    k_iou_loss_fwd<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pred, pred_stride, target, target_stride, n,
                                                                                 loss24);
    return (int)cudaGetLastError();
}

extern "C" int p24_iou_loss_bwd(const float* pred, int64_t pred_stride, const float* target, int64_t target_stride,
                                const float* grad_loss24, int n, float* grad_pred, void* stream) {
    if (n < 0) return P24_E_BADARG;
    if (n == 0) return 0;
    if (!pred || !target || !grad_loss24 || !grad_pred) return P24_E_BADARG;
    k_iou_loss_bwd<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(pred, pred_stride, target, target_stride, grad_loss24, n,
                                                                 grad_pred);
    return (int)cudaGetLastError();
}

extern "C" int p24_pair_iou(const float* gt50, int64_t gt_stride, int G, const float* pred26, int64_t pred_stride, int P,
                            float* out, void* stream) {
    if (G < 0 || P < 0) return P24_E_BADARG;
    if (G == 0 || P == 0) return 0;
    if (!gt50 || !pred26 || !out) return P24_E_BADARG;
    const long long tot = (long long)G * P;
    k_pair_iou<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gt50, gt_stride, G, pred26, pred_stride, P, out);
    return (int)cudaGetLastError();
}

extern "C" int p24_loss_bwd(const float* outputs, int64_t img_stride, int64_t row_stride, int B, int A, int num_classes,
                            const float* labels, int64_t lab_img_stride, int64_t lab_row_stride, const uint8_t* fg_mask,
                            const int32_t* matched_gt, const float* pred_iou, const float* weights_n27,
                            const float* grad_scale, float* grad_outputs, void* stream) {
    if (!outputs || !labels || !fg_mask || !matched_gt || !pred_iou || !weights_n27 || !grad_outputs) return P24_E_BADARG;
    if (B <= 0 || A <= 0 || num_classes <= 0) return P24_E_BADARG;
    const long long rows = (long long)B * A;
    k_loss_bwd<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(outputs, img_stride, row_stride, B, A, num_classes,
                                                                          labels, lab_img_stride, lab_row_stride, fg_mask,
                                                                          matched_gt, pred_iou, weights_n27, grad_scale,
                                                                          grad_outputs);
    return (int)cudaGetLastError();
}

extern "C" int p24_loss_bwd_raw(const float* const* h_raw, const int64_t* h_raw_batch_stride, float* const* h_grad_raw,
                                const int32_t* h_levels, int n_levels, int B, int A, int num_classes, const float* labels,
                                int64_t lab_img_stride, int64_t lab_row_stride, const uint8_t* fg_mask,
                                const int32_t* matched_gt, const float* pred_iou, const float* weights_n27,
                                const float* grad_scale, void* stream) {
    if (!h_raw || !h_raw_batch_stride || !h_grad_raw || !h_levels || !labels || !fg_mask || !matched_gt || !pred_iou ||
        !weights_n27)
        return P24_E_BADARG;
    if (B <= 0 || A <= 0 || num_classes <= 0 || n_levels < 1 || n_levels > 4 || B > 65535) return P24_E_BADARG;
    RawBwd r;
    memset(&r, 0, sizeof(r));
    r.nlev = n_levels;
    int total = 0;
    for (int l = 0; l < n_levels; ++l) {
        const int off = h_levels[4 * l], W = h_levels[4 * l + 1], H = h_levels[4 * l + 2];
        if (off != total || W <= 0 || H <= 0) return P24_E_BADARG;
        r.off[l] = off;
        r.W[l] = W;
        memcpy(&r.st[l], &h_levels[4 * l + 3], 4);
        total += W * H;
        for (int t = 0; t < 3; ++t) {
            r.in[4 * t + l] = h_raw[t * n_levels + l];
            r.in_bs[4 * t + l] = h_raw_batch_stride[t * n_levels + l];
            r.out[4 * t + l] = h_grad_raw[t * n_levels + l];
            if (!r.in[4 * t + l] || !r.out[4 * t + l]) return P24_E_BADARG;
        }
    }
    if (total != A) return P24_E_BADARG;
    r.off[n_levels] = A;
    k_loss_bwd_raw<<<dim3((unsigned)((A + 255) / 256), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
        r, B, A, num_classes, labels, lab_img_stride, lab_row_stride, fg_mask, matched_gt, pred_iou, weights_n27, grad_scale);
    return (int)cudaGetLastError();
}

extern "C" int p24_pack_labels(const double* targets, const int32_t* offsets, const int32_t* shapes_hw, int in_h, int in_w,
                               int B, int max_labels, float* labels, int32_t* nlabel, void* stream) {
    if (!offsets || !shapes_hw || !labels || B <= 0 || max_labels <= 0 || in_h <= 0 || in_w <= 0 || B > 65535)
        return P24_E_BADARG;
    // (targets may be NULL when no image of the batch has a target: every offset is 0)
    const int per = max_labels * 51;
    k_pack_labels<<<dim3((unsigned)((per + 255) / 256), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
        targets, offsets, shapes_hw, in_h, in_w, B, max_labels, labels, nlabel);
    return (int)cudaGetLastError();
}

extern "C" size_t p24_dynamic_k_workspace_bytes(int G, int P) {
    if (G <= 0 || P <= 0) return 0;
    return p24_align((size_t)P * sizeof(int)) + p24_align((size_t)G * P);
}

extern "C" int p24_dynamic_k_matching(const float* cost, const float* ious, int G, int P, uint8_t* fg_in, int32_t* matched,
                                      float* matched_iou, int32_t* dyn_k, int32_t* num_fg, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    if (!cost || !ious || !fg_in || !matched || !matched_iou || !dyn_k || !num_fg || !workspace) return P24_E_BADARG;
    if (G <= 0 || P <= 0) return P24_E_BADARG;
    if (workspace_bytes < p24_dynamic_k_workspace_bytes(G, P)) return P24_E_WORKSPACE;
    int* cnt = (int*)workspace;
    unsigned char* taken = (unsigned char*)workspace + p24_align((size_t)P * sizeof(int));
    cudaStream_t st = (cudaStream_t)stream;
    k_zero_i32<<<(P + 255) / 256, 256, 0, st>>>(cnt, P, num_fg);
    k_dynk_select<<<G, P24_THREADS, 0, st>>>(cost, ious, G, P, cnt, matched, dyn_k, taken);
    k_dynk_resolve<<<(P + 255) / 256, 256, 0, st>>>(cost, ious, G, P, cnt, fg_in, matched, matched_iou, num_fg);
    return (int)cudaGetLastError();
}
