// Host build of the scalar arithmetic of csrc/p24_math.cuh (test infrastructure only: the product never calls the host
// versions).  Compiled by tests/test_hostmath_cpu.py with  g++ -O2 -ffp-contract=off  so that, as on the device
// (-fmad=false), every operation rounds once in the reference's order.
#include "../../exploration-of-potential_b200/csrc/p24_math.cuh"

extern "C" {

void hm_ray_loss(int n, const float* rg, const float* rp, const float* d, float* loss, float* inter) {
    for (int i = 0; i < n; ++i) loss[i] = p24_ray_loss(rg[i], rp[i], d[i], inter + i);
}

// pair value of n (GT, prediction) pairs: rg, rp are [n][24]
void hm_pair_value(int n, const float* rg, const float* rp, const float* d, float* out) {
    for (int i = 0; i < n; ++i) out[i] = p24_pair_value(rg + 24 * i, rp + 24 * i, d[i]);
}

// unsigned angle sum (degrees) of the 24-gon (vx, vy) seen from n points
void hm_angle_sum(int n, const float* vx, const float* vy, const float* x, const float* y, float* out) {
    for (int i = 0; i < n; ++i) out[i] = p24_angle_sum(vx, vy, x[i], y[i]);
}

void hm_in_centre(int n, const float* gcx, const float* gcy, const float* xs, const float* ys, float stride, int* out) {
    for (int i = 0; i < n; ++i)
        out[i] = p24_in_centre(gcx[i], gcy[i], p24_anchor_centre(xs[i], stride), p24_anchor_centre(ys[i], stride), stride) ? 1 : 0;
}

void hm_bce_logits(int n, const float* x, const float* t, float* out) {
    for (int i = 0; i < n; ++i) out[i] = p24_bce_logits(x[i], t[i]);
}

// SimOTA cost of n pairs from the class logits [n][nc], objectness logit, GT class, pair value
void hm_cost(int n, int nc, const float* cls, const float* obj, const int* gt_cls, const float* value, const int* valid,
             float* out) {
    for (int i = 0; i < n; ++i) {
        const float so = p24_sigmoid(obj[i]);
        float s = 0.0f;
        for (int j = 0; j < nc; ++j) {
            const float p = p24_joint_prob(cls[(long long)i * nc + j], so);
            s = s + (j == gt_cls[i] ? p24_bce_pos(p) : p24_bce_neg(p));
        }
        out[i] = p24_cost(s, value[i], valid[i] != 0);
    }
}
}
