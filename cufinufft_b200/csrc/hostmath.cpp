// hostmath.cpp -- plan-time host arithmetic: kernel width/beta from tol, fine-grid
// size, Gauss-Legendre nodes and the phihat quadrature precomputation.
//
// These tiny routines must reproduce the reference's numbers exactly (SURVEY.md
// A.2: a "more accurate" fp32 phihat fails parity), so each keeps the reference's
// evaluation precision step by step:
//   setup_spreader   contrib/spreadinterp.cpp:6-67
//   next235beven     contrib/utils.cpp:3-22
//   SET_NF_TYPE12    contrib/common.cpp:24-37
//   fseries precomp  contrib/common.cpp:84-96 (+ legendre_compute_glr for the nodes)
#include <cmath>
#include <complex>
#include <cstdio>
#include "cfb_plan.h"

namespace cfb {

template <typename T> struct eps_of;
template <> struct eps_of<float>  { static constexpr float  v = (float)6e-08; };   // contrib/utils_fp.h:37
template <> struct eps_of<double> { static constexpr double v = 1.1e-16; };        // contrib/utils_fp.h:45

template <typename T>
int setup_spreader(T eps, double upsampfac_d, int kerevalmeth, int *ns_out, T *beta, T *halfwidth, T *c)
{
    const T PI = (T)M_PI;
    const T upsampfac = (T)upsampfac_d;
    if (upsampfac != 2.0) {
        if (kerevalmeth == 1) {
            fprintf(stderr, "[cufinufft-b200] upsampfac=%.3g cannot be handled by kerevalmeth=1\n", upsampfac_d);
            return 8;
        }
        if (upsampfac <= 1.0) {
            fprintf(stderr, "[cufinufft-b200] upsampfac=%.3g is <= 1.0\n", upsampfac_d);
            return 7;
        }
    }
    int ier = 0;
    if (eps < eps_of<T>::v) {
        fprintf(stderr, "[cufinufft-b200] warning: increasing tol=%.3g to eps_mach=%.3g\n", (double)eps,
                (double)eps_of<T>::v);
        eps = eps_of<T>::v;
        ier = 1;
    }
    int ns = (int)std::ceil(-std::log10(eps / (T)10.0));      // T-precision log10, as the reference's overload
    if (upsampfac != 2.0) ns = (int)std::ceil(-std::log(eps) / (PI * std::sqrt(1 - 1 / upsampfac)));
    ns = ns < 2 ? 2 : ns;
    if (ns > MAX_NS) { ns = MAX_NS; ier = 1; }
    *ns_out = ns;
    *halfwidth = (T)ns / 2;
    *c = (T)(4.0 / (T)(ns * ns));
    T betaoverns = (T)2.30;
    if (ns == 2) betaoverns = (T)2.20;
    if (ns == 3) betaoverns = (T)2.26;
    if (ns == 4) betaoverns = (T)2.38;
    if (upsampfac != 2.0) {
        T gamma = (T)0.97;
        betaoverns = gamma * PI * (1 - 1 / (2 * upsampfac));
    }
    *beta = betaoverns * (T)ns;
    return ier;
}
template int setup_spreader<float>(float, double, int, int *, float *, float *, float *);
template int setup_spreader<double>(double, double, int, int *, double *, double *, double *);

int next235beven(int n, int b)
{
    if (n <= 2) return 2;
    n += n & 1;
    for (int cand = n;; cand += 2) {
        int r = cand;
        for (int p : {2, 3, 5})
            while (r % p == 0) r /= p;
        if (r == 1 && cand % b == 0) return cand;
    }
}

int set_nf_type12(int m, double upsampfac, int ns, int gpu_method, int obinsize)
{
    int nf = (int)(upsampfac * m);
    if (nf < 2 * ns) nf = 2 * ns;
    return next235beven(nf, gpu_method == 4 ? obinsize : 1);
}

// Gauss-Legendre rule of order n on [-1,1], nodes ascending (the convention of the
// reference's legendre_compute_glr).  Newton on the three-term recurrence.
void gauss_legendre(int n, double *x, double *w)
{
    auto eval = [n](double t, double &p, double &dp) {
        double p0 = 1.0, p1 = t;
        for (int k = 2; k <= n; ++k) {
            double p2 = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k;
            p0 = p1;
            p1 = p2;
        }
        p = p1;
        dp = n * (t * p1 - p0) / (t * t - 1.0);
    };
    for (int i = 0; i < (n + 1) / 2; ++i) {
        double t = std::cos(M_PI * (i + 0.75) / (n + 0.5)), p, dp;
        for (int it = 0; it < 100; ++it) {
            eval(t, p, dp);
            double dt = p / dp;
            t -= dt;
            if (std::fabs(dt) < 1e-16) break;
        }
        eval(t, p, dp);
        x[i] = -t;
        x[n - 1 - i] = t;
        w[i] = w[n - 1 - i] = 2.0 / ((1.0 - t * t) * dp * dp);
    }
}

// host ES kernel, contrib/spreadinterp.cpp:69-81: sqrt/exp in double, result narrowed to T
template <typename T>
static T es_kernel_host(T x, T beta, T es_c, T halfwidth)
{
    if (std::abs(x) >= halfwidth) return (T)0.0;
    return (T)std::exp(beta * std::sqrt(1.0 - es_c * x * x));
}

template <typename T>
void fseries_precomp(int nf, int ns, T beta, T es_c, T halfwidth, T *f, double *a_reim)
{
    const T PI = (T)M_PI;
    const std::complex<T> IMA(0.0, 1.0);
    T J2 = (T)(ns / 2.0);
    int q = (int)(2 + 3.0 * J2);
    double z[2 * MAX_NQUAD], w[2 * MAX_NQUAD];
    gauss_legendre(2 * q, z, w);
    for (int n = 0; n < q; ++n) {
        z[n] *= J2;
        f[n] = J2 * (T)w[n] * es_kernel_host<T>((T)z[n], beta, es_c, halfwidth);
        // phase winding rate, evaluated in complex<T> exactly like the reference (A.2 item 7)
        std::complex<T> a = std::exp((T)2 * PI * IMA * (T)(nf / 2 - z[n]) / (T)nf);
        a_reim[2 * n] = (double)a.real();
        a_reim[2 * n + 1] = (double)a.imag();
    }
}
template void fseries_precomp<float>(int, int, float, float, float, float *, double *);
template void fseries_precomp<double>(int, int, double, double, double, double *, double *);

}  // namespace cfb
