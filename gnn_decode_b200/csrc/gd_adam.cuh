// Adam update of one fp32 master weight, the arithmetic of torch.optim.Adam's default (non-amsgrad, L2 weight decay
// folded into the gradient) path -- the optimizer the reference constructs at quantum/decoder_v2_4.py:323
// (Adam(lr=3e-4, weight_decay=1e-9)) and steps at :338.  Shared by gd_adam_step and the fused all-reduce + Adam kernel.
#pragma once
#include "gd_common.cuh"

namespace gd {

struct AdamCoef {
    float beta1, beta2, one_m_beta1, one_m_beta2, eps, weight_decay;
    float step_size;        // lr / (1 - beta1^step)
    float bc2_sqrt;         // sqrt(1 - beta2^step)
};

__host__ inline AdamCoef adam_coef(double lr, double beta1, double beta2, double eps, double weight_decay, int step) {
    AdamCoef c;
    c.beta1 = (float)beta1; c.beta2 = (float)beta2; c.one_m_beta1 = (float)(1.0 - beta1); c.one_m_beta2 = (float)(1.0 - beta2);
    c.eps = (float)eps; c.weight_decay = (float)weight_decay;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    c.step_size = (float)(lr / bc1);
    c.bc2_sqrt = (float)sqrt(bc2);
    return c;
}

__device__ __forceinline__ void adam_update(const AdamCoef& c, float g, float& w, float& m, float& v) {
    g = fmaf(c.weight_decay, w, g);                       // grad.add(param, alpha=weight_decay)
    m = fmaf(c.one_m_beta1, g - m, m);                    // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(c.one_m_beta2 * g, g, c.beta2 * v);          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = __fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt) + c.eps;
    w = fmaf(-c.step_size, __fdiv_rn(m, denom), w);       // param.addcdiv_(exp_avg, denom, value=-step_size)
}

}  // namespace gd
