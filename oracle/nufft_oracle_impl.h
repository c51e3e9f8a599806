/*
 * oracle/nufft_oracle_impl.h  --  TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Precision-generic body of the CPU oracle: a plain-C restatement of the
 * arithmetic of cuFINUFFT v1.3's type-1/type-2 hot path.  Included twice by
 * nufft_oracle.c, once with FLT=double (suffix "") and once with FLT=float
 * (suffix "f").  Every function cites the reference file:line it follows
 * (paths relative to /root/reference).
 *
 * The reference has no CPU spreader (contrib/spreadinterp.cpp holds only
 * setup_spreader + evaluate_kernel), so spread/interp/sort below follow the
 * arithmetic of the reference's *device* kernels, including where it rounds to
 * FLT and where it computes in double.
 */

#define ORC_CAT_(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT_(a, b)
#define ORC(name) ORC_CAT(name, ORC_SUF)

/* ---- kernel parameters: contrib/spreadinterp.cpp:6-67 ------------------- */
int ORC(orc_setup_spreader)(FLT eps, double upsampfac, int kerevalmeth,
                            int *ns_out, FLT *beta_out, FLT *halfwidth_out, FLT *c_out)
{
    FLT ups = (FLT)upsampfac; /* the reference narrows the double option to FLT at the call */
    if (ups != 2.0) {
        if (kerevalmeth == 1) return 8;   /* HORNER_WRONG_BETA, contrib/utils.h:35 */
        if (ups <= 1.0) return 7;         /* ERR_UPSAMPFAC_TOO_SMALL */
    }
    int ier = 0;
    if (eps < ORC_EPSILON) { eps = ORC_EPSILON; ier = 1; }
    int ns = (int)ceil(-ORC_LOG10(eps / (FLT)10.0));          /* spreadinterp.cpp:44; std::log10(FLT) overload */
    if (ups != 2.0)
        ns = (int)ceil(-ORC_LOG(eps) / (ORC_PI * ORC_SQRT(1 - 1 / ups)));
    if (ns < 2) ns = 2;
    if (ns > 16) { ns = 16; ier = 1; }
    *ns_out = ns;
    *halfwidth_out = (FLT)ns / 2;
    *c_out = (FLT)(4.0 / (FLT)(ns * ns));                      /* :56 */
    FLT betaoverns = (FLT)2.30;
    if (ns == 2) betaoverns = (FLT)2.20;
    if (ns == 3) betaoverns = (FLT)2.26;
    if (ns == 4) betaoverns = (FLT)2.38;
    if (ups != 2.0) {
        FLT gamma = (FLT)0.97;
        betaoverns = gamma * ORC_PI * (1 - 1 / (2 * ups));
    }
    *beta_out = betaoverns * (FLT)ns;                          /* :66 */
    return ier;
}

/* host evaluate_kernel: contrib/spreadinterp.cpp:69-81 (1.0 literal => double sqrt/exp) */
static FLT ORC(host_kernel)(FLT x, FLT beta, FLT c, FLT halfwidth)
{
    if (fabs((double)x) >= (double)halfwidth) return (FLT)0.0;
    FLT cx2 = c * x * x;
    return (FLT)exp((double)beta * sqrt(1.0 - (double)cx2));
}

/* device evaluate_kernel: src/cuspreadinterp.h:6-16; es_c, es_beta arrive as
 * double parameters of eval_kernel_vec (:33-40) but es_c*x*x is FLT*FLT in
 * evaluate_kernel itself (its parameters are FLT). */
static inline FLT ORC(dev_kernel)(FLT x, FLT es_c, FLT es_beta, int ns)
{
    FLT ax = (FLT)fabs((double)x);
    if (!((double)ax < ns / 2.0)) return (FLT)0.0;
    FLT cx2 = es_c * ax * ax;
    return (FLT)exp((double)es_beta * sqrt(1.0 - (double)cx2));
}

/* RESCALE macro, pirange=1: contrib/spreadinterp.h:36-38.  M_1_2PI and the
 * 0.5/1.5 literals are double, so the sum and the *N are double; the result is
 * narrowed to FLT by the assignment at every call site. */
static inline FLT ORC(rescale)(FLT x, int nf)
{
    const FLT pi = ORC_PI;
    double shift = (x < -pi) ? 1.5 : ((x >= pi) ? -0.5 : 0.5);
    return (FLT)(((double)x * 0.159154943091895336 + shift) * (double)nf);
}

/* ---- phihat host precomputation: contrib/common.cpp:84-96 ---------------- */
/* z,w: Gauss-Legendre nodes/weights of order 2q on [-1,1] (ascending), as
 * returned by legendre_compute_glr (contrib/legendre_rule_fast.c:15-92).     */
void ORC(orc_fseries_precomp)(int nf, int ns, FLT beta, FLT es_c, FLT halfwidth,
                              FLT *f, double *a_reim /* 2*q doubles */)
{
    FLT J2 = (FLT)(ns / 2.0);
    int q = (int)(2 + 3.0 * J2);
    double z[2 * ORC_MAX_NQUAD], w[2 * ORC_MAX_NQUAD];
    orc_gauss_legendre(2 * q, z, w);
    for (int n = 0; n < q; ++n) {
        z[n] *= J2;
        f[n] = J2 * (FLT)w[n] * ORC(host_kernel)((FLT)z[n], beta, es_c, halfwidth);
        /* a[n] = exp(2*PI*IMA*(FLT)(nf/2-z[n])/(FLT)nf) -- evaluated in complex<FLT>:
         * 2*PI is FLT, times IMA (0,1) => (0, 2PI); times (FLT)(nf/2 - z) then /(FLT)nf. */
        FLT t = (FLT)((double)(nf / 2) - z[n]);
        FLT im = (FLT)((FLT)2 * ORC_PI);
        im = im * t;
        im = im / (FLT)nf;
        /* complex exp of purely imaginary argument in FLT precision */
        a_reim[2 * n + 0] = (double)ORC_COS(im);
        a_reim[2 * n + 1] = (double)ORC_SIN(im);
    }
}

/* ---- phihat device sum: src/common.cu:16-45 ------------------------------ */
void ORC(orc_fseries_compute)(int nf, int ns, const FLT *f, const double *a_reim, FLT *fwkerhalf)
{
    FLT J2 = (FLT)(ns / 2.0);
    int q = (int)(2 + 3.0 * J2);
    for (int i = 0; i < nf / 2 + 1; ++i) {
        int brk = (int)(0.5 + i);
        FLT x = (FLT)0.0;
        for (int n = 0; n < q; ++n) {
            double re = a_reim[2 * n], im = a_reim[2 * n + 1];
            double mag = hypot(re, im);          /* cuCabs */
            double ang = atan2(im, re);          /* carg   */
            x = (FLT)((double)x + (double)f[n] * (2 * (pow(mag, (double)brk) * cos(brk * ang))));
        }
        fwkerhalf[i] = x;
    }
}

/* ---- the CPU phihat of the reference: contrib/common.cpp:98-124 ---------- */
void ORC(orc_fseries_cpu)(int nf, int ns, const FLT *f, const double *a_reim, FLT *fwkerhalf)
{
    FLT J2 = (FLT)(ns / 2.0);
    int q = (int)(2 + 3.0 * J2);
    double ajr[ORC_MAX_NQUAD], aji[ORC_MAX_NQUAD];
    for (int n = 0; n < q; ++n) { ajr[n] = 1.0; aji[n] = 0.0; }
    for (int j = 0; j < nf / 2 + 1; ++j) {
        FLT x = (FLT)0.0;
        for (int n = 0; n < q; ++n) {
            x = (FLT)((double)x + (double)f[n] * 2 * ajr[n]);
            double r = ajr[n] * a_reim[2 * n] - aji[n] * a_reim[2 * n + 1];
            double i = ajr[n] * a_reim[2 * n + 1] + aji[n] * a_reim[2 * n];
            ajr[n] = r; aji[n] = i;
        }
        fwkerhalf[j] = x;
    }
}

/* ---- bin index of one point: src/2d/spreadinterp2d.cu:111-119, 3d:24-39 --- */
static inline int ORC(bin_of)(FLT xr, int binsize, int nbin)
{
    int b = (int)floor(xr / (FLT)binsize);   /* FLT / (int->FLT), floor in FLT */
    b = b >= nbin ? b - 1 : b;
    b = b < 0 ? 0 : b;
    return b;
}

/* setpts bin sort + subproblem map.
 *   CalcBinSize_noghost_*  src/{1,2,3}d/spreadinterp*.cu (1d:75, 2d:103, 3d:16)
 *   exclusive scan         src/2d/spread2d_wrapper.cu:485-487
 *   CalcInvertofGlobalSortIdx_* (2d:129, 3d:45)
 *   CalcSubProb_*          src/precision_independent.cu:66-73 (integer ceil-div here;
 *                          identical to the float ceil for binsize < 2^24)
 *   inclusive scan + MapBintoSubProb_* (:76-85)
 * idxnupts is produced in stable (input) order inside each bin; the reference's
 * within-bin order is a race and only comparable as a per-bin set.
 * Returns totalnumsubprob. subprob_to_bin must hold >= nbins + M/maxsub entries. */
int ORC(orc_binsort)(int dim, int M, const FLT *x, const FLT *y, const FLT *z,
                     int nf1, int nf2, int nf3, int bs1, int bs2, int bs3, int maxsubprobsize,
                     int *binsize, int *binstartpts, int *idxnupts,
                     int *numsubprob, int *subprobstartpts, int *subprob_to_bin)
{
    int nb1 = (int)ceil((FLT)nf1 / bs1);
    int nb2 = dim > 1 ? (int)ceil((FLT)nf2 / bs2) : 1;
    int nb3 = dim > 2 ? (int)ceil((FLT)nf3 / bs3) : 1;
    int nbins = nb1 * nb2 * nb3;
    int *binof = (int *)malloc(sizeof(int) * (size_t)(M > 0 ? M : 1));
    memset(binsize, 0, sizeof(int) * (size_t)nbins);
    for (int i = 0; i < M; ++i) {
        int b = ORC(bin_of)(ORC(rescale)(x[i], nf1), bs1, nb1);
        if (dim > 1) b += nb1 * ORC(bin_of)(ORC(rescale)(y[i], nf2), bs2, nb2);
        if (dim > 2) b += nb1 * nb2 * ORC(bin_of)(ORC(rescale)(z[i], nf3), bs3, nb3);
        binof[i] = b;
        binsize[b]++;
    }
    int acc = 0;
    for (int b = 0; b < nbins; ++b) { binstartpts[b] = acc; acc += binsize[b]; }
    int *fill = (int *)calloc((size_t)nbins, sizeof(int));
    for (int i = 0; i < M; ++i) { int b = binof[i]; idxnupts[binstartpts[b] + fill[b]++] = i; }
    free(fill); free(binof);
    subprobstartpts[0] = 0;
    int T = 0;
    for (int b = 0; b < nbins; ++b) {
        numsubprob[b] = (binsize[b] + maxsubprobsize - 1) / maxsubprobsize;
        for (int s = 0; s < numsubprob[b]; ++s) subprob_to_bin[T + s] = b;
        T += numsubprob[b];
        subprobstartpts[b + 1] = T;
    }
    return T;
}

/* ---- kernel vectors ------------------------------------------------------ */
/* eval_kernel_vec: src/cuspreadinterp.h:33-40 */
static inline void ORC(kervec)(FLT *ker, FLT x1, int ns, FLT es_c, FLT es_beta)
{
    for (int i = 0; i < ns; ++i)
        ker[i] = ORC(dev_kernel)((FLT)fabs((double)(x1 + (FLT)i)), es_c, es_beta, ns);
}

/* eval_kernel_vec_Horner: src/cuspreadinterp.h:18-31 with the reference's table
 * (contrib/ker_horner_allw_loop.c:4-216, imported into oracle/horner_ref_table.inc).
 * z = 2x + w - 1.0 is double then FLT. */
static inline void ORC(kervec_horner)(FLT *ker, FLT x, int w)
{
    FLT zz = (FLT)(2 * (double)x + w - 1.0);
    const double *tab = orc_horner_table(w);
    int nc = orc_horner_ncoef(w);
    for (int i = 0; i < w; ++i) {
        FLT acc = (FLT)tab[(nc - 1) * 16 + i];
        for (int k = nc - 2; k >= 0; --k) acc = (FLT)tab[k * 16 + i] + zz * acc;
        ker[i] = acc;
    }
}

/* exported views for the pinning tests (tests/test_oracle_pinned.py) */
void ORC(orc_horner_eval)(int w, FLT x1, FLT *ker) { ORC(kervec_horner)(ker, x1, w); }
void ORC(orc_host_kernel_vec)(int n, const FLT *x, FLT beta, FLT c, FLT halfwidth, FLT *out)
{
    for (int i = 0; i < n; ++i) out[i] = ORC(host_kernel)(x[i], beta, c, halfwidth);
}

static inline int ORC(wrap)(int i, int nf) { return i < 0 ? i + nf : (i > nf - 1 ? i - nf : i); }

/* stencil start + kernel vector for one coordinate:
 * src/2d/spreadinterp2d.cu:35-43 (NUptsdriven) / :192-199 (Subprob) */
static inline int ORC(stencil)(FLT xr, int ns, FLT es_c, FLT es_beta, int horner, FLT *ker)
{
    int xstart = (int)ceil((double)xr - ns / 2.0);
    FLT x1 = (FLT)xstart - xr;
    if (horner) ORC(kervec_horner)(ker, x1, ns);
    else        ORC(kervec)(ker, x1, ns, es_c, es_beta);
    return xstart;
}

/* ---- spread (type-1 step 1): Spread_{1,2,3}d_NUptsdriven[_Horner]
 * src/1d/spreadinterp1d.cu:17, src/2d/spreadinterp2d.cu:17-98, src/3d/spreadinterp3d.cu:74-175.
 * The stencil is exactly ns points from xstart (the reference's
 * occasional (ns+1)-th point reads an uninitialised weight whose intended value is 0).
 * fw is [nf3][nf2][nf1] complex interleaved, NOT zeroed here. Accumulation is in FLT,
 * product order cnow*k1*k2*k3 as in the reference.  OpenMP: atomic adds. */
void ORC(orc_spread)(int dim, long M, const FLT *x, const FLT *y, const FLT *z, const FLT *c,
                     int nf1, int nf2, int nf3, int ns, FLT es_c, FLT es_beta, int horner, FLT *fw)
{
#pragma omp parallel for schedule(static)
    for (long j = 0; j < M; ++j) {
        FLT k1[16], k2[16], k3[16];
        int xs = ORC(stencil)(ORC(rescale)(x[j], nf1), ns, es_c, es_beta, horner, k1);
        int ys = 0, zs = 0;
        if (dim > 1) ys = ORC(stencil)(ORC(rescale)(y[j], nf2), ns, es_c, es_beta, horner, k2);
        if (dim > 2) zs = ORC(stencil)(ORC(rescale)(z[j], nf3), ns, es_c, es_beta, horner, k3);
        FLT cr = c[2 * j], ci = c[2 * j + 1];
        int nz = dim > 2 ? ns : 1, ny = dim > 1 ? ns : 1;
        for (int iz = 0; iz < nz; ++iz) {
            int gz = dim > 2 ? ORC(wrap)(zs + iz, nf3) : 0;
            for (int iy = 0; iy < ny; ++iy) {
                int gy = dim > 1 ? ORC(wrap)(ys + iy, nf2) : 0;
                for (int ix = 0; ix < ns; ++ix) {
                    int gx = ORC(wrap)(xs + ix, nf1);
                    size_t o = ((size_t)gz * nf2 + gy) * nf1 + gx;
                    FLT vr = cr * k1[ix], vi = ci * k1[ix];
                    if (dim > 1) { vr *= k2[iy]; vi *= k2[iy]; }
                    if (dim > 2) { vr *= k3[iz]; vi *= k3[iz]; }
#pragma omp atomic
                    fw[2 * o] += vr;
#pragma omp atomic
                    fw[2 * o + 1] += vi;
                }
            }
        }
    }
}

/* ---- interp (type-2 step 3): Interp_{1,2,3}d_NUptsdriven[_Horner]
 * src/1d/spreadinterp1d.cu:234, src/2d/spreadinterp2d.cu:517-593, src/3d/spreadinterp3d.cu:655-756 */
void ORC(orc_interp)(int dim, long M, const FLT *x, const FLT *y, const FLT *z, FLT *c,
                     int nf1, int nf2, int nf3, int ns, FLT es_c, FLT es_beta, int horner, const FLT *fw)
{
#pragma omp parallel for schedule(static)
    for (long j = 0; j < M; ++j) {
        FLT k1[16], k2[16], k3[16];
        int xs = ORC(stencil)(ORC(rescale)(x[j], nf1), ns, es_c, es_beta, horner, k1);
        int ys = 0, zs = 0;
        if (dim > 1) ys = ORC(stencil)(ORC(rescale)(y[j], nf2), ns, es_c, es_beta, horner, k2);
        if (dim > 2) zs = ORC(stencil)(ORC(rescale)(z[j], nf3), ns, es_c, es_beta, horner, k3);
        FLT cr = 0, ci = 0;
        int nz = dim > 2 ? ns : 1, ny = dim > 1 ? ns : 1;
        for (int iz = 0; iz < nz; ++iz) {
            int gz = dim > 2 ? ORC(wrap)(zs + iz, nf3) : 0;
            for (int iy = 0; iy < ny; ++iy) {
                int gy = dim > 1 ? ORC(wrap)(ys + iy, nf2) : 0;
                for (int ix = 0; ix < ns; ++ix) {
                    int gx = ORC(wrap)(xs + ix, nf1);
                    size_t o = ((size_t)gz * nf2 + gy) * nf1 + gx;
                    FLT wgt = k1[ix];
                    if (dim > 1) wgt *= k2[iy];
                    if (dim > 2) wgt *= k3[iz];
                    cr += fw[2 * o] * wgt;
                    ci += fw[2 * o + 1] * wgt;
                }
            }
        }
        c[2 * j] = cr; c[2 * j + 1] = ci;
    }
}

/* ---- deconvolve / amplify: src/deconvolve_wrapper.cu:14-121 --------------- */
/* dir=1: fk[i] = fw[w]/(ker product);  dir=2: fw[w] = fk[i]/(ker product) (fw pre-zeroed by caller,
 * as the reference memsets it: src/deconvolve_wrapper.cu:141,181,228). */
void ORC(orc_deconvolve)(int dir, int dim, int ms, int mt, int mu, int nf1, int nf2, int nf3,
                         FLT *fw, FLT *fk, const FLT *ker1, const FLT *ker2, const FLT *ker3)
{
    if (dim < 2) mt = 1;
    if (dim < 3) mu = 1;
#pragma omp parallel for schedule(static)
    for (int k3 = 0; k3 < mu; ++k3) {
        int w3 = dim > 2 ? (k3 - mu / 2 >= 0 ? k3 - mu / 2 : nf3 + k3 - mu / 2) : 0;
        for (int k2 = 0; k2 < mt; ++k2) {
            int w2 = dim > 1 ? (k2 - mt / 2 >= 0 ? k2 - mt / 2 : nf2 + k2 - mt / 2) : 0;
            for (int k1 = 0; k1 < ms; ++k1) {
                int w1 = k1 - ms / 2 >= 0 ? k1 - ms / 2 : nf1 + k1 - ms / 2;
                size_t in = ((size_t)w3 * nf2 + w2) * nf1 + w1;
                size_t out = ((size_t)k3 * mt + k2) * ms + k1;
                FLT kv = ker1[abs(k1 - ms / 2)];
                if (dim > 1) kv = kv * ker2[abs(k2 - mt / 2)];
                if (dim > 2) kv = kv * ker3[abs(k3 - mu / 2)];
                if (dir == 1) { fk[2 * out] = fw[2 * in] / kv; fk[2 * out + 1] = fw[2 * in + 1] / kv; }
                else          { fw[2 * in] = fk[2 * out] / kv; fw[2 * in + 1] = fk[2 * out + 1] / kv; }
            }
        }
    }
}

/* ---- direct sums in double at sampled outputs: semantics of contrib/dirft2d.cpp:7-74
 * (and the 3-D loops of test/cufinufft3d1_test.cu:175-184), k from -m/2 .. (m-1)/2. */
void ORC(orc_dirft1_sampled)(int dim, long M, const FLT *x, const FLT *y, const FLT *z, const FLT *c,
                             int iflag, int ms, int mt, int mu, int nsamp, const long *modeidx, double *out)
{
    if (dim < 2) mt = 1;
    if (dim < 3) mu = 1;
    double sgn = iflag >= 0 ? 1.0 : -1.0;
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < nsamp; ++s) {
        long i = modeidx[s];
        int i1 = (int)(i % ms), i2 = (int)((i / ms) % mt), i3 = (int)(i / ms / mt);
        double k1 = i1 - ms / 2, k2 = dim > 1 ? i2 - mt / 2 : 0, k3 = dim > 2 ? i3 - mu / 2 : 0;
        double sr = 0, si = 0;
        for (long j = 0; j < M; ++j) {
            double ph = k1 * (double)x[j];
            if (dim > 1) ph += k2 * (double)y[j];
            if (dim > 2) ph += k3 * (double)z[j];
            double cs = cos(ph), sn = sgn * sin(ph);
            double cr = c[2 * j], ci = c[2 * j + 1];
            sr += cr * cs - ci * sn;
            si += cr * sn + ci * cs;
        }
        out[2 * s] = sr; out[2 * s + 1] = si;
    }
}

void ORC(orc_dirft2_sampled)(int dim, const FLT *x, const FLT *y, const FLT *z, const FLT *fk,
                             int iflag, int ms, int mt, int mu, int nsamp, const long *ptidx, double *out)
{
    if (dim < 2) mt = 1;
    if (dim < 3) mu = 1;
    double sgn = iflag >= 0 ? 1.0 : -1.0;
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < nsamp; ++s) {
        long j = ptidx[s];
        double sr = 0, si = 0;
        for (int i3 = 0; i3 < mu; ++i3)
            for (int i2 = 0; i2 < mt; ++i2)
                for (int i1 = 0; i1 < ms; ++i1) {
                    double ph = (i1 - ms / 2) * (double)x[j];
                    if (dim > 1) ph += (i2 - mt / 2) * (double)y[j];
                    if (dim > 2) ph += (i3 - mu / 2) * (double)z[j];
                    double cs = cos(ph), sn = sgn * sin(ph);
                    size_t o = ((size_t)i3 * mt + i2) * ms + i1;
                    double fr = fk[2 * o], fi = fk[2 * o + 1];
                    sr += fr * cs - fi * sn;
                    si += fr * sn + fi * cs;
                }
        out[2 * s] = sr; out[2 * s + 1] = si;
    }
}

#undef ORC
#undef ORC_CAT
#undef ORC_CAT_
