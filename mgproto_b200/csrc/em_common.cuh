// Shared by em.cu and em_tc.cu: Adam configuration, the closed-form zero-gradient replay, one Adam step.
#pragma once
#include <math.h>
#include "mgp_common.cuh"

namespace mgp_em {

constexpr float EM_EPS = 1e-10f;

// ---------------------------------------------------------------------------------------------
// torch.optim.Adam takes its hyper-parameters as Python doubles and derives 1 - beta in DOUBLE before the value is
// narrowed to the tensor's fp32 (`lerp_(grad, 1 - beta1)`, `addcmul_(grad, grad, value=1 - beta2)`): 1.0f - 0.999f
// would be off by 1.3e-5 relative.  The C ABI therefore carries them as doubles and the fp32 factors are derived here.
struct AdamCfg {
    double lr, beta1, beta2, eps;
    double ln_b1, ln_b2;            // log(beta) evaluated on the host: the planner's tables are exp(t ln b), not pow(b, t)
    float b2f, omb1, omb2, epsf;
};
static inline AdamCfg make_adam(double lr, double beta1, double beta2, double eps) {
    AdamCfg a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    a.ln_b1 = (beta1 > 0.0) ? log(beta1) : -1e300; a.ln_b2 = (beta2 > 0.0) ? log(beta2) : -1e300;
    a.b2f = (float)beta2; a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2); a.epsf = (float)eps;
    return a;
}

// ---------------------------------------------------------------------------------------------
// Zero-gradient Adam steps in closed form.  With g = 0 the moments just decay (m_s = b1^s m_0, v_s = b2^s v_0) and
//   p <- p - sum_s c_s m_0 / (sqrt(v_0) d_s + eps),   c_s = lr b1^s / (1 - b1^(first+s)),
//                                                     d_s = b2^(s/2) / sqrt(1 - b2^(first+s)).
// The first S = ceil(log 1e-3 / log b1) steps (66 for b1 = 0.9; 1e-6 -> 132 while first < 2000) are applied term by term from a table of (c_s, d_s);
// the remaining steps S+1 .. count carry b1^s < 1e-3 of the first term's weight and are folded into ONE term: their
// c_s form a geometric series (the bias correction 1 - b1^(first+s) is constant to 1e-3 * 1e-3 there) and d_s, which
// drifts by b2^(1/2) per step, is taken at the series' centre of mass s* = S + 1 + b1/(1-b1).  Against the exact
// step-by-step recursion this is accurate to < 4e-7 of the replay's total movement (tools/em_replay_check.py), below
// the fp32 rounding of the reference's own step-by-step evaluation.
__device__ __forceinline__ int replay_explicit_steps(int count, int first, float beta1) {
    int cutoff = count;
    // a young optimiser (first < 2000) keeps terms down to 1e-6: there 1 - b2^t still moves by > 1e-3 per step
    if (beta1 > 0.f && beta1 < 1.f) cutoff = (int)ceilf(logf(first >= 2000 ? 1e-3f : 1e-6f) / logf(beta1));
    return min(count, max(cutoff, 1));
}
__device__ __forceinline__ void replay_coeffs(const AdamCfg& a, int first, int s, float& c, float& d) {
    const double st = (double)s;
    const double l1 = log(a.beta1), l2 = log(a.beta2);
    const double b1s = exp(st * l1), b2s = exp(st * l2);
    const double b1t = exp(((double)first + st) * l1);
    const double b2t = exp(((double)first + st) * l2);
    c = (float)(a.lr * b1s / (1.0 - b1t));
    d = (float)(sqrt(b2s) / sqrt(1.0 - b2t));
}
__device__ __forceinline__ void replay_tail(const AdamCfg& a, int first, int S, int count, float& c, float& d) {
    const double b1 = a.beta1, l1 = log(b1), l2 = log(a.beta2);
    const double geo = exp((S + 1.0) * l1) * (1.0 - exp((double)(count - S) * l1)) / (1.0 - b1);
    c = (float)(a.lr * geo / (1.0 - exp(((double)first + S + 1.0) * l1)));
    const int ss = min(count, S + 1 + (int)(b1 / (1.0 - b1)));
    d = (float)(sqrt(exp((double)ss * l2)) / sqrt(1.0 - exp(((double)first + ss) * l2)));
}

__device__ __forceinline__ void adam_apply(float& p, float& m, float& v, float g, const AdamCfg& a, double b1pow,
                                           double b2pow) {
    // torch.optim.Adam (_single_tensor_adam): lerp, mul/addcmul, bias corrections in double
    m = m + (g - m) * a.omb1;
    v = v * a.b2f + a.omb2 * g * g;
    const float step_size = (float)(a.lr / (1.0 - b1pow));
    const float bc2_sqrt = (float)sqrt(1.0 - b2pow);
    const float denom = sqrtf(v) / bc2_sqrt + a.epsf;
    p = p - step_size * (m / denom);
}


}  // namespace mgp_em
