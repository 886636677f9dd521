// TEST INFRASTRUCTURE.  Host build of nflows_b200/csrc/rq_spline.cuh: the exact source the CUDA kernels evaluate per
// element, compiled for the CPU (ex2.approx -> exp2f), exported through a tiny C ABI so tests can study the numerics of
// the kernel formulation against the fp64 oracle without a GPU.  Built by oracle/Makefile into oracle/_build/.
#include <string.h>

#include "../nflows_b200/csrc/rq_spline.cuh"

namespace nfk {
int make_spline_params_host(const NfkSplineDesc* d, SplineParams* p) {
    const int K = d->num_bins;
    p->num_bins = K;
    p->linear_tails = d->linear_tails ? 1 : 0;
    p->left = (float)d->left; p->right = (float)d->right; p->bottom = (float)d->bottom; p->top = (float)d->top;
    p->span_w = (float)(d->right - d->left);
    p->span_h = (float)(d->top - d->bottom);
    p->min_w = (float)d->min_bin_width; p->min_h = (float)d->min_bin_height; p->min_d = (float)d->min_derivative;
    p->mix_w = (float)(1.0 - d->min_bin_width * K);
    p->mix_h = (float)(1.0 - d->min_bin_height * K);
    p->beta = (float)d->softplus_beta;
    p->inv_beta = (float)(1.0 / d->softplus_beta);
    float div = (float)d->wh_divisor;
    p->pre_scale = (d->wh_divisor == 1.0) ? 1.0f : (float)(1.0 / (double)div);
    p->edge_ud = (float)log(exp(1.0 - d->min_derivative) - 1.0);
    p->knot_eps = 1e-6f;
    return 0;
}
}  // namespace nfk

extern "C" int rqs_host_eval(const NfkSplineDesc* desc, int inverse, const float* x, const float* uw, const float* uh,
                             const float* ud, long long n, float* y, float* lad, int* flags) {
    constexpr int KMAX = 16;
    nfk::SplineParams p;
    nfk::make_spline_params_host(desc, &p);
    const int K = p.num_bins;
    if (K > KMAX) return -1;
    const int nd = p.linear_tails ? K - 1 : K + 1;
    int flag = 0;
    for (long long e = 0; e < n; ++e) {
        float w[KMAX] = {0}, h[KMAX] = {0}, d[KMAX + 1] = {0};
        for (int k = 0; k < K; ++k) { w[k] = uw[e * K + k]; h[k] = uh[e * K + k]; }
        if (p.linear_tails) {
            for (int k = 0; k <= KMAX; ++k) d[k] = (k >= 1 && k < K) ? ud[e * nd + k - 1] : p.edge_ud;
        } else {
            for (int k = 0; k <= K; ++k) d[k] = ud[e * nd + k];
        }
        nfk::rqs_eval<KMAX>(p, inverse != 0, x[e], w, h, d, y[e], lad[e], flag);
    }
    if (flags) *flags = flag;
    return 0;
}
