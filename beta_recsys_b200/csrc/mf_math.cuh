// Per-sample MF maths shared by the MF kernels: score -> loss -> d loss / d score.
//   MF.forward          beta_rec/models/mf.py:32-55   (sigmoid BEFORE the BPR difference)
//   bpr_loss / bce_loss beta_rec/models/torch_engine.py:92-121
#pragma once
#include "common.cuh"

enum { LOSS_BPR = 0, LOSS_BCE = 1 };

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_scale(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 b) {
    return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w));
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// zp / zn: pre-sigmoid scores of the (u,i) and (u,j) pairs (dot + biases); rating: BCE target.
// Returns the sample's loss term and cu_i = d(mean loss)/d zp, cu_j = d(mean loss)/d zn.
template <int LOSS>
__device__ __forceinline__ void mf_sample_coef(float zp, float zn, float rating, float inv_b, float& cu_i, float& cu_j,
                                               float& loss_k) {
    if (LOSS == LOSS_BPR) {
        // mf.py:43-48 then torch_engine.py:104-105
        const float sp = sigmoidf_(zp);
        const float sn = sigmoidf_(zn);
        const float d = sp - sn;
        loss_k = -logsigmoidf_(d);
        const float dx = -inv_b / (1.0f + expf(d));  // d/dx of -mean(logsigmoid(x))
        cu_i = dx * sp * (1.0f - sp);
        cu_j = -dx * sn * (1.0f - sn);
    } else {
        // nn.BCELoss: logs clamped at -100; backward (s-r)/max((1-s)s, 1e-12)/B
        const float sc_ = sigmoidf_(zp);
        loss_k = -(rating * fmaxf(logf(sc_), -100.f) + (1.0f - rating) * fmaxf(log1pf(-sc_), -100.f));
        const float ds = (sc_ - rating) / fmaxf((1.0f - sc_) * sc_, 1e-12f) * inv_b;
        cu_i = ds * sc_ * (1.0f - sc_);
        cu_j = 0.f;
    }
}

// Same quantities with the SFU intrinsics (ex2 / lg2 / rcp approximations, relative error ~1e-6) for the BPR
// chain of the row-owner kernels, where the chain's ~120 instructions per block were a visible share of the
// issue slots.  Every argument stays in a benign range (|d| < 1, 1 + e in (1, 2]), so the results differ from
// the ATen-style evaluation above by ~1e-7 absolute -- two orders below the 1e-5 parity budget.  BCE keeps the
// precise path: its log(1 - s) is ill-conditioned for saturated scores.
template <int LOSS>
__device__ __forceinline__ void mf_sample_coef_fast(float zp, float zn, float rating, float inv_b, float& cu_i,
                                                    float& cu_j, float& loss_k) {
    if (LOSS == LOSS_BPR) {
        const float sp = __fdividef(1.0f, 1.0f + __expf(-zp));
        const float sn = __fdividef(1.0f, 1.0f + __expf(-zn));
        const float d = sp - sn;
        const float e = __expf(-fabsf(d));
        const float t = __fdividef(1.0f, 1.0f + e);  // sigmoid(|d|)
        loss_k = __logf(1.0f + e) - fminf(d, 0.0f);  // -logsigmoid(d)
        const float dx = -inv_b * (d < 0.f ? t : e * t);  // -sigmoid(-d) / B
        cu_i = dx * sp * (1.0f - sp);
        cu_j = -dx * sn * (1.0f - sn);
    } else {
        mf_sample_coef<LOSS>(zp, zn, rating, inv_b, cu_i, cu_j, loss_k);
    }
}
