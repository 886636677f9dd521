// Rational-quadratic spline, one element per thread, everything in registers.
//
// Numerical contract = reference nflows/transforms/splines/rational_quadratic.py:13-181
// and nflows/utils/torchutils.py:134-136, restated for a thread that owns all K unnormalised widths/heights
// and the K+1 unnormalised derivatives of ONE element:
//   * softmax with max-subtraction and a true division, min-size mixing, sequential prefix sum, affine map to
//     [lo,hi], first/last knot forced (:91-98, :106-113);
//   * bin = last k in [0,K-1] with x >= knot_k (the reference counts x >= knot over all K+1 knots after adding
//     1e-6 to the last one; identical for in-domain x because knots are non-decreasing);
//   * only the two derivatives the bin needs go through softplus (threshold 20, like F.softplus);
//   * forward: theta=(x-cw)/w ...; inverse: root = 2c / (-b - sqrt(b^2-4ac)) (:132-181).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nfk.h"

namespace nfk {

struct SplineParams {        // derived on the host in double precision, rounded once to fp32 (like the
    int num_bins;            // Python-float -> fp32-scalar conversions the reference's tensor ops perform)
    int linear_tails;
    float left, right, bottom, top;
    float span_w, span_h;    // float(right-left), float(top-bottom)
    float min_w, min_h, min_d;
    float mix_w, mix_h;      // float(1 - min_w*K), float(1 - min_h*K)
    float beta, inv_beta;    // softplus beta
    float pre_scale;         // float(1/float(sqrt(H))): ATen divides by a CPU scalar as x * (1/s)
    float edge_ud;           // float(log(exp(1-min_d)-1)): boundary unnormalised derivative for linear tails
    float knot_eps;          // 1e-6f, searchsorted eps
};

// host: derive the fp32 constants from the descriptor (nfk_spline.cu)
int make_spline_params(const NfkSplineDesc* d, SplineParams* p);

__device__ __forceinline__ float softplus_torch(float x, float beta, float inv_beta) {
    // F.softplus(x, beta, threshold=20): x*beta > 20 ? x : log1p(exp(x*beta))/beta
    float xb = x * beta;
    return xb > 20.0f ? x : log1pf(expf(xb)) * inv_beta;
}

// KMAX: compile-time bound on the number of bins (register arrays); p.num_bins <= KMAX.
// uw/uh: K raw values (before the 1/sqrt(H) pre-scale); ud: K+1 raw derivative logits (already padded with
// edge_ud for linear tails).  Returns y and log|dy/dx|; sets flag bits for domain / discriminant violations.
template <int KMAX>
__device__ __forceinline__ void rqs_eval(const SplineParams& p, bool inverse, float x, const float (&uw)[KMAX],
                                         const float (&uh)[KMAX], const float (&ud)[KMAX + 1], float& y, float& lad,
                                         int& flag) {
    const int K = p.num_bins;
    if (p.linear_tails) {
        bool inside = (x >= p.left) && (x <= p.right);   // NaN -> outside -> identity, lad 0 (reference :26-39)
        if (!inside) { y = x; lad = 0.0f; return; }
    } else if (!(x >= p.left && x <= p.right)) {
        flag |= 1;                                        // reference raises InputOutsideDomain (:81-82)
        x = fminf(fmaxf(x, p.left), p.right);
    }

    float ew[KMAX], eh[KMAX];
    float mw = -INFINITY, mh = -INFINITY;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
            ew[k] = uw[k] * p.pre_scale;
            eh[k] = uh[k] * p.pre_scale;
            mw = fmaxf(mw, ew[k]);
            mh = fmaxf(mh, eh[k]);
        }
    }
    float sw = 0.0f, sh = 0.0f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
            ew[k] = expf(ew[k] - mw);
            eh[k] = expf(eh[k] - mh);
            sw += ew[k];
            sh += eh[k];
        }
    }

    // running prefix sums -> knots; select the bin on the fly
    const float q = x;
    float cum_w = 0.0f, cum_h = 0.0f;
    float kw_lo = p.left, kh_lo = p.bottom;              // knot k
    float b_cw = p.left, b_ch = p.bottom, b_w = 1.0f, b_h = 1.0f;
    int bin = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
            float fw = p.min_w + p.mix_w * (ew[k] / sw);
            float fh = p.min_h + p.mix_h * (eh[k] / sh);
            cum_w += fw;
            cum_h += fh;
            float kw_hi = (k == K - 1) ? p.right : p.span_w * cum_w + p.left;   // knot k+1
            float kh_hi = (k == K - 1) ? p.top : p.span_h * cum_h + p.bottom;
            bool take = (k == 0) || (q >= (inverse ? kh_lo : kw_lo));
            if (take) {
                bin = k;
                b_cw = kw_lo; b_ch = kh_lo;
                b_w = kw_hi - kw_lo; b_h = kh_hi - kh_lo;
            }
            kw_lo = kw_hi; kh_lo = kh_hi;
        }
    }

    float ud0 = ud[0], ud1 = ud[1];
#pragma unroll
    for (int k = 1; k < KMAX; ++k) {
        if (k == bin) { ud0 = ud[k]; ud1 = ud[k + 1]; }
    }
    const float d0 = p.min_d + softplus_torch(ud0, p.beta, p.inv_beta);
    const float d1 = p.min_d + softplus_torch(ud1, p.beta, p.inv_beta);
    const float delta = b_h / b_w;
    const float s = d0 + d1 - 2.0f * delta;

    float theta;
    if (inverse) {
        float u = x - b_ch;
        float a = u * s + b_h * (delta - d0);
        float b = b_h * d0 - u * s;
        float c = -delta * u;
        float disc = b * b - 4.0f * a * c;
        if (!(disc >= 0.0f)) flag |= 2;                   // reference: assert (discriminant >= 0).all() (:142)
        theta = (2.0f * c) / (-b - sqrtf(disc));
        y = theta * b_w + b_cw;
    } else {
        theta = (x - b_cw) / b_w;
    }
    const float t1mt = theta * (1.0f - theta);
    const float den = delta + s * t1mt;
    if (!inverse) {
        float num = b_h * (delta * (theta * theta) + d0 * t1mt);
        y = b_ch + num / den;
    }
    const float omt = 1.0f - theta;
    const float dnum = (delta * delta) * (d1 * (theta * theta) + 2.0f * delta * t1mt + d0 * (omt * omt));
    const float l = logf(dnum) - 2.0f * logf(den);
    lad = inverse ? -l : l;
}

}  // namespace nfk
