// oracle/ref_gpu_shim.cu -- TEST INFRASTRUCTURE.  Our own extern "C" window into the
// REFERENCE library's public plan struct (include/cufinufft_eitherprec.h:247-297), compiled
// against /root/reference headers and linked with the reference objects into
// oracle/_ref/libcufinufft_ref.so.  Lets the GPU parity tests read the reference's bin
// counts / offsets / subproblem map / phihat and run its spread-only and interp-only
// stages (CUSPREADnD / CUINTERPnD, src/cuspreadinterp.h:276-339) on a fully made plan.
// Compiled twice (double, -DSINGLE).
#include <cufinufft_eitherprec.h>
#include "../src/cuspreadinterp.h"

#ifdef SINGLE
#define SFX(n) n##f
#else
#define SFX(n) n
#endif

extern "C" {

// what: 0 geometry {dim,nf1,nf2,nf3,ns,nbins1,nbins2,nbins3,bsx,bsy,bsz,maxbatch,M,totalnumsubprob,method,nbins}
//       1 binsize 2 binstartpts 3 numsubprob 4 subprobstartpts 5 subprob_to_bin 6 idxnupts
int SFX(refg_get_ints)(CUFINUFFT_PLAN p, int what, int *out)
{
    int dim = p->dim;
    int bs[3] = {p->opts.gpu_binsizex, dim > 1 ? p->opts.gpu_binsizey : 1, dim > 2 ? p->opts.gpu_binsizez : 1};
    int nf[3] = {p->nf1, p->nf2, p->nf3};
    int nb[3] = {1, 1, 1}, nbins = 1;
    for (int d = 0; d < dim; ++d) { nb[d] = ceil((FLT)nf[d] / bs[d]); nbins *= nb[d]; }
    cudaDeviceSynchronize();
    switch (what) {
        case 0: {
            int g[16] = {dim, p->nf1, p->nf2, p->nf3, p->spopts.nspread, nb[0], nb[1], nb[2], bs[0], bs[1], bs[2],
                         p->maxbatchsize, p->M, p->totalnumsubprob, p->opts.gpu_method, nbins};
            for (int i = 0; i < 16; ++i) out[i] = g[i];
            return 0;
        }
        case 1: return cudaMemcpy(out, p->binsize, nbins * sizeof(int), cudaMemcpyDeviceToHost);
        case 2: return cudaMemcpy(out, p->binstartpts, nbins * sizeof(int), cudaMemcpyDeviceToHost);
        case 3: return cudaMemcpy(out, p->numsubprob, nbins * sizeof(int), cudaMemcpyDeviceToHost);
        case 4: return cudaMemcpy(out, p->subprobstartpts, (nbins + 1) * sizeof(int), cudaMemcpyDeviceToHost);
        case 5: return cudaMemcpy(out, p->subprob_to_bin, p->totalnumsubprob * sizeof(int), cudaMemcpyDeviceToHost);
        case 6: return cudaMemcpy(out, p->idxnupts, p->M * sizeof(int), cudaMemcpyDeviceToHost);
    }
    return 1;
}

int SFX(refg_get_reals)(CUFINUFFT_PLAN p, int d, FLT *out)
{
    cudaDeviceSynchronize();
    if (d == -1) { out[0] = p->spopts.ES_beta; out[1] = p->spopts.ES_c; out[2] = p->spopts.ES_halfwidth; return 0; }
    FLT *src = d == 0 ? p->fwkerhalf1 : (d == 1 ? p->fwkerhalf2 : p->fwkerhalf3);
    int nf = d == 0 ? p->nf1 : (d == 1 ? p->nf2 : p->nf3);
    return cudaMemcpy(out, src, (nf / 2 + 1) * sizeof(FLT), cudaMemcpyDeviceToHost);
}

// spread-only / interp-only of ONE transform with the plan's own method; fw must be nf1*nf2*nf3
int SFX(refg_spread)(CUFINUFFT_PLAN p, CUCPX *c, CUCPX *fw)
{
    CUCPX *sc = p->c, *sfw = p->fw;
    p->c = c; p->fw = fw;
    cudaMemset(fw, 0, (size_t)p->nf1 * p->nf2 * p->nf3 * sizeof(CUCPX));
    int ier = p->dim == 1 ? CUSPREAD1D(p, 1) : (p->dim == 2 ? CUSPREAD2D(p, 1) : CUSPREAD3D(p, 1));
    cudaDeviceSynchronize();
    p->c = sc; p->fw = sfw;
    return ier;
}
int SFX(refg_interp)(CUFINUFFT_PLAN p, CUCPX *c, CUCPX *fw)
{
    CUCPX *sc = p->c, *sfw = p->fw;
    p->c = c; p->fw = fw;
    int ier = p->dim == 1 ? CUINTERP1D(p, 1) : (p->dim == 2 ? CUINTERP2D(p, 1) : CUINTERP3D(p, 1));
    cudaDeviceSynchronize();
    p->c = sc; p->fw = sfw;
    return ier;
}
}
