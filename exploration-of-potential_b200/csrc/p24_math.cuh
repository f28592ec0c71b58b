// p24_math.cuh — scalar fp32 arithmetic of the YOLOX-24p loss / SimOTA path.
//
// Every expression here follows the operation order of SURVEY.md Appendix A, i.e. the order in
// which the reference's eager PyTorch code rounds (one rounding per op, no FMA contraction).
// The translation unit that includes this header for the device MUST be compiled with
// `-fmad=false` (and the default -prec-div=true -prec-sqrt=true): discrete SimOTA decisions hang
// on fp32 thresholds, so the association order is part of the contract.
//
// The functions are `__host__ __device__` so that a host-only build (tests/tools/hostmath.cpp,
// test infrastructure) can check the formulas against the oracle on a CPU-only machine; the
// product never calls the host versions.
//
// Reference lines (paths relative to /root/reference/yolox_24p):
//   ray term            models/losses.py:36-72,118-151  ==  utils/boxes.py:127-157,201-235
//   polygon angle sum   models/losses.py:566-588
//   centre window       models/losses.py:523-542
//   class cost          models/losses.py:399-416
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define P24_HD __host__ __device__ __forceinline__
#else
#define P24_HD inline
#endif

#define P24_RAYS 24

// fp32 constants exactly as torch materialises them
#define P24_PI 3.1415927410125732f        // torch.tensor(np.pi)
#define P24_RAD2DEG 57.295780181884766f   // torch.rad2deg multiplier
#define P24_PENALTY 100000.0f

// ---------------------------------------------------------------------------------------------
// One ray of the concentric-circle GIoU: returns loss24 = 1 - giou.  `inter_out` (optional)
// receives the intersection area that IOUloss.circle_inter returns.
// ---------------------------------------------------------------------------------------------
P24_HD float p24_ray_loss(float rg, float rp, float d, float* inter_out = nullptr) {
    const float rmin = fminf(rg, rp);
    const float rmax = fmaxf(rg, rp);
    const float rmin2 = rmin * rmin;
    const float rmax2 = rmax * rmax;
    const bool nested = fabsf(rg - rp) >= d;   // losses.py:60  (written first)
    const bool apart = d >= rg + rp;           // losses.py:67  (written second: overrides)
    float inter;
    if (apart) {
        inter = 0.0f;
    } else if (nested) {
        inter = P24_PI * rmin2;
    } else {
        const float d2 = d * d;
        float ac_min = ((rmin2 + d2) - rmax2) / (((2.0f * rmin) * d) + 1e-8f);
        float ac_max = ((rmax2 + d2) - rmin2) / (((2.0f * rmax) * d) + 1e-8f);
        ac_min = fminf(fmaxf(ac_min, -0.99f), 0.99f);
        ac_max = fminf(fmaxf(ac_max, -0.99f), 0.99f);
        const float ang_min = acosf(ac_min);
        const float ang_max = acosf(ac_max);
        inter = ((ang_min * rmin2) + (ang_max * rmax2)) - ((rmin * d) * sinf(ang_min));
    }
    if (inter_out) *inter_out = inter;
    const float ag = P24_PI * (rg * rg);
    const float ap = P24_PI * (rp * rp);
    const float uni = (ag + ap) - inter;
    // (apart: inter is +0 and uni + 1e-6 is positive or NaN, which reaches giou through uni anyway)
    const float iou = apart ? 0.0f : inter / (uni + 1e-6f);
    const float cl = nested ? rmax : (((rg + rp) + d) / 2.0f);
    const float cs = P24_PI * (cl * cl);
    const float giou = iou - ((cs - uni) / cs);
    return 1.0f - giou;
}

// centre distance, losses.py:36 / boxes.py:127
P24_HD float p24_centre_dist(float gcx, float gcy, float pcx, float pcy) {
    const float dx = gcx - pcx;
    const float dy = gcy - pcy;
    return sqrtf((dx * dx) + (dy * dy));
}

// GT ray length: torch.norm over the 2-vector (losses.py:108, boxes.py:197).  ATen's CUDA norm
// reduction accumulates `acc + x*x`, which nvcc contracts to an FMA; the same form is used here.
P24_HD float p24_gt_radius(float vx, float vy) {
    return sqrtf(fmaf(vy, vy, vx * vx));
}

// pairwise "iou" of SimOTA: mean ray loss / 2 (boxes.py:238-241).  rp points at 24 radii.
P24_HD float p24_pair_value(const float* rg, const float* rp, float d) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < P24_RAYS; ++k) s = s + p24_ray_loss(rg[k], rp[k], d);
    return (s / 24.0f) / 2.0f;
}

// ---------------------------------------------------------------------------------------------
// polygon "inside" test: total unsigned angular variation >= 350 degrees (losses.py:583-588)
// vx, vy: 24 vertices.  Returns the angle sum in degrees.
// ---------------------------------------------------------------------------------------------
P24_HD float p24_angle_sum(const float* vx, const float* vy, float xc, float yc) {
    float acc = 0.0f;
    float sx = vx[0] - xc, sy = vy[0] - yc;
    const float fx = sx, fy = sy;
#pragma unroll
    for (int k = 0; k < P24_RAYS; ++k) {
        const float ex = (k == P24_RAYS - 1) ? fx : (vx[k + 1] - xc);
        const float ey = (k == P24_RAYS - 1) ? fy : (vy[k + 1] - yc);
        const float cross = (sx * ey) - (ex * sy);
        const float dot = (sx * ex) + (sy * ey);
        acc = acc + (atan2f(fabsf(cross), dot) * P24_RAD2DEG);
        sx = ex;
        sy = ey;
    }
    return acc;
}

// centre window, strict (losses.py:523-542)
P24_HD bool p24_in_centre(float gcx, float gcy, float xc, float yc, float stride) {
    const float r = 2.5f * stride;
    const float cl = xc - (gcx - r);
    const float cr = (gcx + r) - xc;
    const float ct = yc - (gcy - r);
    const float cb = (gcy + r) - yc;
    return fminf(fminf(cl, ct), fminf(cr, cb)) > 0.0f;
}

// anchor centre (losses.py:506-516)
P24_HD float p24_anchor_centre(float shift, float stride) {
    return (shift * stride) + (0.5f * stride);
}

// ---------------------------------------------------------------------------------------------
// class cost terms (losses.py:409-416): p = sqrt(sigmoid(cls) * sigmoid(obj)),
// BCE(p, y) = -(y * max(log p, -100) + (1 - y) * max(log1p(-p), -100))
// ---------------------------------------------------------------------------------------------
P24_HD float p24_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

P24_HD float p24_joint_prob(float cls_logit, float obj_sig) { return sqrtf(p24_sigmoid(cls_logit) * obj_sig); }

P24_HD float p24_bce_neg(float p) { return -fmaxf(log1pf(-p), -100.0f); }  // target 0
P24_HD float p24_bce_pos(float p) { return -fmaxf(logf(p), -100.0f); }     // target 1

// full SimOTA cost (losses.py:420-424)
P24_HD float p24_cost(float cls_cost, float pair_value, bool valid) {
    const float iou_cost = -logf(pair_value + 1e-8f);
    return (cls_cost + (3.0f * iou_cost)) + (valid ? 0.0f : P24_PENALTY);
}

// BCEWithLogits (losses.py:294-302): max(x,0) - x*t + log1p(exp(-|x|))  (ATen's stable form)
P24_HD float p24_bce_logits(float x, float t) {
    return ((1.0f - t) * x) + (fmaxf(-x, 0.0f) + log1pf(expf(-fabsf(x))));
}


#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// Fast polygon test (device only).  The reference thresholds the total UNSIGNED angle subtended by the
// 24 edges at 350 degrees (losses.py:583-588).  Instead of 24 atan2 the unsigned angles are accumulated as
// the argument of a complex product  prod_k (dot_k + i |cross_k|)  (each factor has argument in [0, pi]),
// counting wraps across 2 pi.  Returns 1 (sum >= 350.05 deg), 0 (sum <= 349.95 deg) or 2 (within 0.05 deg of
// the threshold, or degenerate): only then is the reference-order p24_angle_sum evaluated.  The
// 0.05 degree band is > 50x the fp32 error of either evaluation (< 1e-3 degrees).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int p24_angle_test_fast(const float* __restrict__ vx, const float* __restrict__ vy, float xc,
                                                   float yc) {
    float re = 1.0f, im = 0.0f;
    int wraps = 0;
    float sx = vx[0] - xc, sy = vy[0] - yc;
    const float fx = sx, fy = sy;
#pragma unroll
    for (int k = 0; k < P24_RAYS; ++k) {
        const float ex = (k == P24_RAYS - 1) ? fx : (vx[k + 1] - xc);
        const float ey = (k == P24_RAYS - 1) ? fy : (vy[k + 1] - yc);
        const float cr = fabsf(fmaf(sx, ey, -(ex * sy)));
        const float dt = fmaf(sx, ex, sy * ey);
        const float nre = fmaf(re, dt, -(im * cr));
        const float nim = fmaf(re, cr, im * dt);
        wraps += (im < 0.0f && nim >= 0.0f) ? 1 : 0;
        re = nre;
        im = nim;
        sx = ex;
        sy = ey;
        if ((k & 3) == 3) {  // renormalise by a power of two (exact) to stay inside the fp32 range
            const float m = fmaxf(fabsf(re), fabsf(im));
            const int e = (__float_as_int(m) >> 23) & 0xFF;
            const float sc = __int_as_float((254 - e) << 23);
            re *= sc;
            im *= sc;
        }
    }
    if (!(fabsf(re) + fabsf(im) > 0.0f) || !(re == re) || !(im == im) || isinf(re) || isinf(im)) return 2;
    if (wraps > 0) return 1;
    if (im < 0.0f && re > 0.0f) {
        const float y = -im;
        if (y < re * 0.17542f) return 1;   // tan(9.95 deg) = 0.175426
        if (y > re * 0.17723f) return 0;   // tan(10.05 deg) = 0.177226
        return 2;
    }
    return 0;
}

// reference-order evaluation, kept out of line (rare) so that it does not inflate the callers' registers
static __device__ __noinline__ bool p24_in_polygon_exact(const float* vx, const float* vy, float xc, float yc) {
    float acc = 0.0f;
    float sx = vx[0] - xc, sy = vy[0] - yc;
#pragma unroll 1
    for (int k = 0; k < P24_RAYS; ++k) {
        const int k2 = (k == P24_RAYS - 1) ? 0 : k + 1;
        const float ex = vx[k2] - xc, ey = vy[k2] - yc;
        const float cross = (sx * ey) - (ex * sy);
        const float dot = (sx * ex) + (sy * ey);
        acc = acc + (atan2f(fabsf(cross), dot) * P24_RAD2DEG);
        sx = ex;
        sy = ey;
    }
    return acc >= 350.0f;
}

// exact decision, fast path first
__device__ __forceinline__ bool p24_in_polygon(const float* __restrict__ vx, const float* __restrict__ vy, float xc,
                                               float yc) {
    const int r = p24_angle_test_fast(vx, vy, xc, yc);
    if (r != 2) return r == 1;
    return p24_in_polygon_exact(vx, vy, xc, yc);
}

// one edge term of the reference's angle sum (losses.py:572-587), for lane-parallel evaluation
__device__ __forceinline__ float p24_edge_angle(float sx, float sy, float ex, float ey) {
    const float cross = (sx * ey) - (ex * sy);
    const float dot = (sx * ex) + (sy * ey);
    return atan2f(fabsf(cross), dot) * P24_RAD2DEG;
}

// upper bound of one ray's loss in fast arithmetic: exact (to ~1e-6) for apart rays, >= the loss of partial
// rays (2 - uni/cs with the apart formula) and of nested rays (<= 1).  DESIGN.md "top-10 bracket".
__device__ __forceinline__ float p24_ray_loss_ub(float rg, float rp, float d) {
    const float q = (rg + rp) + d;
    return fmaxf(1.0f, 2.0f - __fdividef(4.0f * fmaf(rg, rg, rp * rp), q * q));
}

// ---------------------------------------------------------------------------------------------
// Class cost in product form (device only).  sum_j -log1p(-p_j) = -log prod_j (1 - p_j) with
// p_j = sqrt(sigmoid(cls_j) sigmoid(obj)) = rsqrt((1 + e^-cls_j)(1 + e^-obj)); one log per anchor instead of 80.
// A lane accumulates the factors of its classes; saturated factors (p >= 1, the reference clamps the
// term at 100) are counted separately.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void p24_neg_factor(float cls_logit, float eo1, float& prod, int& nsat) {
    const float q = (1.0f + __expf(-cls_logit)) * eo1;
    const float om = 1.0f - rsqrtf(q);
    if (om > 0.0f) prod *= om;
    else ++nsat;
}
#endif
