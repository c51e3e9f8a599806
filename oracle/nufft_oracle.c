/*
 * oracle/nufft_oracle.c  --  TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the cuFINUFFT v1.3 type-1/type-2 hot path: a plain-C
 * restatement of the reference's algorithm, used ONLY by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * as the checker / CPU baseline.  The product (cufinufft_b200/csrc) never links,
 * loads or calls this file.
 *
 * Parity status: PINNED.  (1) The host math here (setup_spreader, next235beven,
 * nf selection, phihat precomputation and CPU phihat) is checked bit-for-bit
 * against the reference's own contrib/*.cpp compiled from /root/reference into
 * oracle/_ref/libref_host.so (tests/test_oracle_vs_ref_host.py + the committed
 * fixtures in tests/golden/ made by tests/golden/make_golden.py).  (2) The
 * device-path restatements (bin sort, spread, interp, deconvolve) are checked on
 * the GPU box against the reference library itself, built for sm_100 from
 * /root/reference into oracle/_ref/libcufinufft_ref.so (tests/test_vs_reference_gpu.py),
 * and against direct sums (contrib/dirft2d.cpp semantics).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>

#define ORC_MAX_NQUAD 100 /* contrib/common.h:10 */

/* ---- Gauss-Legendre nodes/weights on [-1,1], ascending order --------------
 * Same output convention as legendre_compute_glr (contrib/legendre_rule_fast.c:15-92):
 * x[0] is the most negative node (checked against the built reference).  Newton iteration on P_n with the three-term
 * recurrence; agrees with the reference's Glaser-Liu-Rokhlin values to ~1 ulp. */
void orc_gauss_legendre(int n, double *x, double *w)
{
    for (int i = 0; i < (n + 1) / 2; ++i) {
        double t = cos(M_PI * (i + 0.75) / (n + 0.5));
        double pp = 0;
        for (int it = 0; it < 100; ++it) {
            double p0 = 1.0, p1 = t;
            for (int k = 2; k <= n; ++k) {
                double p2 = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k;
                p0 = p1; p1 = p2;
            }
            pp = n * (t * p1 - p0) / (t * t - 1.0);
            double dt = p1 / pp;
            t -= dt;
            if (fabs(dt) < 1e-16) break;
        }
        /* recompute derivative at the converged node */
        {
            double p0 = 1.0, p1 = t;
            for (int k = 2; k <= n; ++k) {
                double p2 = ((2.0 * k - 1.0) * t * p1 - (k - 1.0) * p0) / k;
                p0 = p1; p1 = p2;
            }
            pp = n * (t * p1 - p0) / (t * t - 1.0);
        }
        x[i] = -t;
        x[n - 1 - i] = t;
        w[i] = w[n - 1 - i] = 2.0 / ((1.0 - t * t) * pp * pp);
    }
}

/* ---- next235beven: contrib/utils.cpp:3-22 --------------------------------- */
int orc_next235beven(int n, int b)
{
    if (n <= 2) return 2;
    if (n % 2 == 1) n += 1;
    int cand = n;
    for (;; cand += 2) {
        int r = cand;
        while (r % 2 == 0) r /= 2;
        while (r % 3 == 0) r /= 3;
        while (r % 5 == 0) r /= 5;
        if (r == 1 && cand % b == 0) return cand;
    }
}

/* ---- SET_NF_TYPE12: contrib/common.cpp:24-37 ------------------------------ */
int orc_set_nf(int ms, double upsampfac, int ns, int gpu_method, int obinsize)
{
    int nf = (int)(upsampfac * ms);
    if (nf < 2 * ns) nf = 2 * ns;
    return orc_next235beven(nf, gpu_method == 4 ? obinsize : 1);
}

/* ---- SETUP_BINSIZE: src/cufinufft.cu:17-73 -------------------------------- */
void orc_default_binsize(int dim, int gpu_method, int *bs /*3*/, int *obs /*3*/)
{
    if (dim == 1) { if (bs[0] < 0) bs[0] = 1024; bs[1] = 1; bs[2] = 1; }
    else if (dim == 2) { if (bs[0] < 0) bs[0] = 32; if (bs[1] < 0) bs[1] = 32; bs[2] = 1; }
    else if (gpu_method == 1 || gpu_method == 2) {
        if (bs[0] < 0) bs[0] = 16; if (bs[1] < 0) bs[1] = 16; if (bs[2] < 0) bs[2] = 2;
    } else if (gpu_method == 4) {
        for (int d = 0; d < 3; ++d) { if (obs[d] < 0) obs[d] = 8; if (bs[d] < 0) bs[d] = 4; }
    }
}

/* ---- Horner tables: the reference's generated constants, contrib/ker_horner_allw_loop.c:4-216 */
#include "horner_ref_table.inc"   /* the oracle's OWN copy of the reference table (tools/import_horner_table.py) */
static const double *orc_horner_table(int w) { return orcref_horner_coeffs[w]; }
static int orc_horner_ncoef(int w) { return orcref_horner_ncoef[w]; }
const double *orc_horner_table_export(int w) { return orc_horner_table(w); }
int orc_horner_ncoef_export(int w) { return orc_horner_ncoef(w); }

/* ---- double precision instantiation --------------------------------------- */
#define FLT double
#define ORC_SUF
#define ORC_EPSILON 1.1e-16            /* contrib/utils_fp.h:45 */
#define ORC_PI ((double)M_PI)
#define ORC_COS cos
#define ORC_SIN sin
#define ORC_LOG10 log10
#define ORC_LOG log
#define ORC_SQRT sqrt
#include "nufft_oracle_impl.h"
#undef FLT
#undef ORC_SUF
#undef ORC_EPSILON
#undef ORC_PI
#undef ORC_COS
#undef ORC_SIN
#undef ORC_LOG10
#undef ORC_LOG
#undef ORC_SQRT

/* ---- single precision instantiation --------------------------------------- */
#define FLT float
#define ORC_SUF f
#define ORC_EPSILON ((float)6e-08)     /* contrib/utils_fp.h:37 */
#define ORC_PI ((float)M_PI)
#define ORC_COS cosf
#define ORC_SIN sinf
#define ORC_LOG10 log10f
#define ORC_LOG logf
#define ORC_SQRT sqrtf
#include "nufft_oracle_impl.h"
