// torch.optim's single-tensor update formulas (SGD / Adam / RMSprop) as the kernels evaluate them --
// shared by rows_apply.cu (scratch-based apply, dense sweep) and mf_rowwise.cu (row-owner MF step).
// Reference: ModelEngine.set_optimizer (beta_rec/models/torch_engine.py:23-39) -> torch/optim/{sgd,adam,rmsprop}.py
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------
// optimizer math
// ---------------------------------------------------------------------------
struct OptScalars {  // per-step scalars, computed in double like torch does on the host
    float lr;
    float one_minus_b1, b2, one_minus_b2;
    float step_size;  // lr / (1 - b1^t)
    float bc2_sqrt;   // sqrt(1 - b2^t)
    float eps;
    float alpha, one_minus_alpha;
};

struct OptParams {
    int kind;
    double lr, beta1, beta2, eps, alpha;
};

__device__ __forceinline__ OptScalars make_scalars(const OptParams& o, long long t) {
    OptScalars s;
    s.lr = (float)o.lr;
    s.one_minus_b1 = (float)(1.0 - o.beta1);
    s.b2 = (float)o.beta2;
    s.one_minus_b2 = (float)(1.0 - o.beta2);
    const double bc1 = 1.0 - pow(o.beta1, (double)t);
    const double bc2 = 1.0 - pow(o.beta2, (double)t);
    s.step_size = (float)(o.lr / bc1);
    s.bc2_sqrt = (float)sqrt(bc2);
    s.eps = (float)o.eps;
    s.alpha = (float)o.alpha;
    s.one_minus_alpha = (float)(1.0 - o.alpha);
    return s;
}

// one element of torch.optim's single-tensor update
template <int KIND>
__device__ __forceinline__ void opt_elem(float& p, float g, float& m, float& v, const OptScalars& s) {
    if (KIND == BRS_SGD) {
        p -= s.lr * g;  // sgd.py: param.add_(grad, alpha=-lr)
    } else if (KIND == BRS_ADAM) {
        m = m + (g - m) * s.one_minus_b1;                // exp_avg.lerp_(grad, 1-beta1)
        v = v * s.b2 + s.one_minus_b2 * g * g;           // mul_(beta2).addcmul_(grad, grad, 1-beta2)
        const float denom = sqrtf(v) / s.bc2_sqrt + s.eps;
        p -= s.step_size * (m / denom);                  // addcdiv_(exp_avg, denom, -step_size)
    } else {
        v = v * s.alpha + s.one_minus_alpha * g * g;     // rmsprop.py, momentum 0, not centered
        p -= s.lr * (g / (sqrtf(v) + s.eps));
    }
}

template <int KIND>
__device__ __forceinline__ void opt_elem4(float4& p, const float4& g, float4& m, float4& v, const OptScalars& s) {
    opt_elem<KIND>(p.x, g.x, m.x, v.x, s);
    opt_elem<KIND>(p.y, g.y, m.y, v.y, s);
    opt_elem<KIND>(p.z, g.z, m.z, v.z, s);
    opt_elem<KIND>(p.w, g.w, m.w, v.w, s);
}

