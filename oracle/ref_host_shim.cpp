// oracle/ref_host_shim.cpp -- TEST INFRASTRUCTURE.  Our own thin extern "C" view of the
// REFERENCE's host math, compiled against /root/reference headers and linked with the
// reference's contrib/*.cpp objects into oracle/_ref/libref_host.so (see Makefile).
// Compiled twice: FLT=double (suffix "") and -DSINGLE FLT=float (suffix "f").
#include <complex>
#include "common.h"      // reference contrib/common.h (FLT, SPREAD_OPTS, onedim_fseries_*)
#include "dirft.h"       // reference contrib/dirft.h
#include <cufinufft_opts.h>

#ifdef SINGLE
#define SFX(n) n##f
#else
#define SFX(n) n
#endif

extern "C" {
#include "legendre_rule_fast.h"
}

extern "C" {

int SFX(refh_setup_spreader)(FLT eps, double upsampfac, int kerevalmeth,
                             int *ns, FLT *beta, FLT *halfwidth, FLT *c)
{
    SPREAD_OPTS o;
    int ier = setup_spreader(o, eps, (FLT)upsampfac, kerevalmeth);
    *ns = o.nspread; *beta = o.ES_beta; *halfwidth = o.ES_halfwidth; *c = o.ES_c;
    return ier;
}

static SPREAD_OPTS SFX(mkopts)(int ns, FLT beta, FLT halfwidth, FLT c)
{
    SPREAD_OPTS o; o.nspread = ns; o.spread_direction = 1; o.pirange = 1; o.upsampfac = 2.0;
    o.ES_beta = beta; o.ES_halfwidth = halfwidth; o.ES_c = c; return o;
}

FLT SFX(refh_evaluate_kernel)(FLT x, int ns, FLT beta, FLT halfwidth, FLT c)
{
    return evaluate_kernel(x, SFX(mkopts)(ns, beta, halfwidth, c));
}

void SFX(refh_fseries_precomp)(int nf, int ns, FLT beta, FLT halfwidth, FLT c, FLT *f, double *a_reim)
{
    dcomplex a[MAX_NQUAD];
    onedim_fseries_kernel_precomp(nf, f, a, SFX(mkopts)(ns, beta, halfwidth, c));
    int q = (int)(2 + 3.0 * (ns / 2.0));
    for (int n = 0; n < q; ++n) { a_reim[2 * n] = a[n].real(); a_reim[2 * n + 1] = a[n].imag(); }
}

void SFX(refh_fseries_cpu)(int nf, int ns, FLT beta, FLT halfwidth, FLT c, FLT *fwkerhalf)
{
    onedim_fseries_kernel(nf, fwkerhalf, SFX(mkopts)(ns, beta, halfwidth, c));
}

void SFX(refh_dirft2d1)(int nj, FLT *x, FLT *y, FLT *c, int iflag, int ms, int mt, FLT *f)
{
    dirft2d1(nj, x, y, (CPX *)c, iflag, ms, mt, (CPX *)f);
}
void SFX(refh_dirft2d2)(int nj, FLT *x, FLT *y, FLT *c, int iflag, int ms, int mt, FLT *f)
{
    dirft2d2(nj, x, y, (CPX *)c, iflag, ms, mt, (CPX *)f);
}

/* the reference's generated piecewise-polynomial table, evaluated on the host exactly as
 * eval_kernel_vec_Horner does on the device (src/cuspreadinterp.h:18-31): the table file is a
 * code fragment that expects `w`, `z` and `ker` in scope. */
void SFX(refh_horner)(int w, FLT x, FLT *ker)
{
    FLT z = 2 * x + w - 1.0;
#include "ker_horner_allw_loop.c"
}

#ifndef SINGLE
int refh_next235beven(int n, int b) { return next235beven(n, b); }
int refh_set_nf(int ms, double upsampfac, int ns, int gpu_method, int obinsize)
{
    cufinufft_opts o; o.upsampfac = upsampfac; o.gpu_method = gpu_method;
    SPREAD_OPTS so; so.nspread = ns;
    BIGINT nf; SET_NF_TYPE12(ms, o, so, &nf, obinsize); return nf;
}
void refh_legendre(int n, double *x, double *w) { legendre_compute_glr(n, x, w); }
#endif
}
