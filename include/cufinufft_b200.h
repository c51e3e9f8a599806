/* include/cufinufft_b200.h -- extension symbols of libcufinufft.so (B200 build).
 *
 * The reference's opts struct cannot grow (see cufinufft_opts.h), so everything
 * the reference does not have goes through these extra extern "C" symbols.
 * None of them is needed for drop-in use.
 *
 *  - stream control + host-pointer convenience calls;
 *  - stage-level entry points (spread / interp only) replacing the reference's
 *    internal, C++-mangled CUFINUFFT_SPREADnD / CUFINUFFT_INTERPnD test hooks
 *    (src/cuspreadinterp.h:278-291, src/2d/spread2d_wrapper.cu:15-97);
 *  - plan introspection replacing direct reads of the reference's public plan
 *    struct (include/cufinufft_eitherprec.h:247-297) by tests: bin counts,
 *    offsets, subproblem map, phihat;
 *  - per-stage device timings (the reference's -DTIME printf instrumentation,
 *    src/2d/cufinufft2d.cu:49-89).
 * Every function returns 0 on success.  "f" suffix = single precision.       */
#ifndef CUFINUFFT_B200_H
#define CUFINUFFT_B200_H

#include "cufinufft.h"

#ifdef __cplusplus
extern "C" {
#endif

/* error codes: 1..9 as the reference (contrib/utils.h:28-36), 10+ are ours */
enum {
    CFB_OK = 0,
    CFB_WARN_EPS_TOO_SMALL = 1,
    CFB_ERR_UPSAMPFAC_TOO_SMALL = 7,
    CFB_ERR_HORNER_WRONG_BETA = 8,
    CFB_ERR_BAD_ARG = 10,       /* NULL plan / bad type, dim, method, bin size ... */
    CFB_ERR_CUDA = 11,          /* a CUDA runtime call failed (cudaGetLastError text on stderr) */
    CFB_ERR_CUFFT = 12,
    CFB_ERR_NOT_IMPLEMENTED = 13, /* type 3 (absent in the reference too, src/cufinufft.cu:533-536) */
    CFB_ERR_NO_POINTS_SET = 14,
    CFB_ERR_NCCL = 15           /* an NCCL call failed (ncclGetErrorString text on stderr) */
};

const char *cufinufft_b200_version(void);

/* Host-only views of the plan-time arithmetic (no CUDA call; usable without a device):
 *  host_params: what makeplan derives from (type, dim, nmodes, tol, opts) --
 *    out_ints16 = {ns, nf1, nf2, nf3, binsx, binsy, binsz, nbins1, nbins2, nbins3, nbins_total,
 *                  engine (1 GM | 2 SM), sorted, maxbatchsize, gpu_method, nquad}
 *    out_reals3 = {ES_beta, ES_c, ES_halfwidth}
 *    (reference: setup_spreader contrib/spreadinterp.cpp:6-67, SET_NF_TYPE12 contrib/common.cpp:24-37,
 *     SETUP_BINSIZE src/cufinufft.cu:17-73); opts may be NULL (defaults).
 *  phihat_quadrature: the f[n] (float or double, by single_precision) and a[n] (re,im doubles)
 *    of onedim_fseries_kernel_precomp, contrib/common.cpp:84-96. */
int cufinufft_b200_host_params(int type, int dim, const int *nmodes, double tol, int single_precision,
                               const cufinufft_opts *opts, int *out_ints16, double *out_reals3);
int cufinufft_b200_phihat_quadrature(int nf, int ns, double beta, double es_c, double halfwidth, int single_precision,
                                     void *f, double *a_reim);
/*  host_workplan: what setpts would choose for M points on a B200 (148 SMs, 227 KB shared memory) --
 *    out_ints16 = {ibinsx, ibinsy, ibinsz (internal bin = reference bin / sub-bins per bin), sub-bins per
 *                  reference bin x, y, z, #internal bins, #reference bins, points per work item, own work
 *                  list (0|1), tile cells, tile stride y, tile stride z, warps per SM-spread block,
 *                  halo ceil(ns/2), bank-conflict cost of the tile layout}
 *    (no reference counterpart: the reference's work items are its bins, src/2d/spread2d_wrapper.cu:386-613). */
int cufinufft_b200_host_workplan(int type, int dim, const int *nmodes, double tol, int single_precision,
                                 const cufinufft_opts *opts, long long M, int *out_ints16);

/* Measured peak of one SM resource on `device`, whole GPU (csrc/microbench.cu; SURVEY.md 8d asks for measured,
 * not nominal, denominators of the binding-resource roofline):
 *   what 0 shared-memory read bandwidth (conflict-free LDS.128), bytes/s     1 FFMA2 (fma.rn.f32x2), FMA/s
 *        2 scalar FFMA, FMA/s      3 DFMA, FMA/s      4 shared-memory atomicAdd(float), conflict-free, atomics/s
 *        5 shared-memory read bandwidth through LDS.64, bytes/s
 * Runs a few ms of synthetic kernels on the device's default stream. */
int cufinufft_b200_microbench(int what, int device, double *out);

/* Stream on which all work of the plan is enqueued (default: the calling thread's per-thread default
 * stream, cudaStreamPerThread, as in the reference build).
 * `stream` is a cudaStream_t passed as void*. */
int cufinufft_set_stream(cufinufft_plan plan, void *stream);
int cufinufftf_set_stream(cufinufftf_plan plan, void *stream);

/* Host-buffer convenience: same as setpts/execute but x,y,z / c,fk are HOST pointers
 * (pinned or pageable); copies are issued on the plan's stream and execute_host returns
 * after the result is back in host memory. */
int cufinufft_setpts_host(int M, const double *x, const double *y, const double *z, cufinufft_plan plan);
int cufinufftf_setpts_host(int M, const float *x, const float *y, const float *z, cufinufftf_plan plan);
int cufinufft_execute_host(cuDoubleComplex *c, cuDoubleComplex *fk, cufinufft_plan plan);
int cufinufftf_execute_host(cuFloatComplex *c, cuFloatComplex *fk, cufinufftf_plan plan);

/* Stage-level calls on a plan that has points set (device pointers):
 *   spread:  fw[ntransf-batch][nf3][nf2][nf1] = sum_j c_j phi(.)   (fw is zeroed first)
 *   interp:  c_j = sum fw phi(.)
 * `nt` transforms are processed (c stride M, fw stride nf1*nf2*nf3). */
int cufinufft_spread(cuDoubleComplex *c, cuDoubleComplex *fw, int nt, cufinufft_plan plan);
int cufinufftf_spread(cuFloatComplex *c, cuFloatComplex *fw, int nt, cufinufftf_plan plan);
int cufinufft_interp(cuDoubleComplex *c, cuDoubleComplex *fw, int nt, cufinufft_plan plan);
int cufinufftf_interp(cuFloatComplex *c, cuFloatComplex *fw, int nt, cufinufftf_plan plan);

/* Introspection.  `what` selects the quantity; ints are written to out[0..]:
 *   geometry (host): 0 -> {dim, nf1, nf2, nf3, ns, nbins1, nbins2, nbins3, binsx, binsy, binsz,
 *                          maxbatchsize, M, totalnumsubprob, method, nbins_total}
 * device arrays copied to the HOST buffer `out` (caller sizes it from the geometry):
 *   1 binsize[nbins]  2 binstartpts[nbins]  3 numsubprob[nbins]  4 subprobstartpts[nbins+1]
 *   5 subprob_to_bin[totalnumsubprob]  6 idxnupts[M]                                          */
int cufinufft_get_ints(cufinufft_plan plan, int what, int *out);
int cufinufftf_get_ints(cufinufftf_plan plan, int what, int *out);
/* fwkerhalf (phihat) of dimension d=0,1,2: nf_d/2+1 reals to HOST buffer out; kernel params:
 * d=-1 -> {beta, c, halfwidth} */
int cufinufft_get_reals(cufinufft_plan plan, int d, double *out);
int cufinufftf_get_reals(cufinufftf_plan plan, int d, float *out);

/* Device time (ms) of the stages of the LAST execute (synchronises the stream):
 * out[0]=spread|interp, out[1]=fft, out[2]=deconvolve|amplify, out[3]=memset, out[4]=total.
 * Timing is recorded only after cufinufft*_set_timing(plan, 1). */
int cufinufft_set_timing(cufinufft_plan plan, int on);
int cufinufftf_set_timing(cufinufftf_plan plan, int on);
int cufinufft_get_timing(cufinufft_plan plan, float *out);
int cufinufftf_get_timing(cufinufftf_plan plan, float *out);
/* Interpolation engine: 0 = automatic (the shared-memory tile engine whenever the points are
 * bin-sorted and the bin tile with its halo fits in shared memory, else the gather engine),
 * 1 = gather engine (reference Interp_*_NUptsdriven schedule), 2 = tile engine when possible
 * (reference Interp_*_Subprob schedule).  Results agree to rounding; this is an A/B switch. */
int cufinufft_set_interp_engine(cufinufft_plan plan, int engine);
int cufinufftf_set_interp_engine(cufinufftf_plan plan, int engine);
/* Order of the points inside a bin (takes effect at the next setpts): 0 = automatic, 1 = one level
 * (global histogram over (bin, stencil cell) keys), 2 = two levels (bins globally, stencil cells per
 * work item in shared memory -- chosen automatically when the one-level histogram would exceed 2 GB).
 * Add 4 to switch the coarse-partition locality pass in front of the counting sort off, 8 to force it on
 * (automatic: on once the point records outgrow a quarter of L2; csrc/setpts.cu: coarse_scatter_kernel).
 * Add 16 to compute the sort keys with the generic code even for power-of-two bin sizes (csrc/setpts.cu: key_any).
 * Results do not depend on it; an A/B and test switch. */
int cufinufft_set_sort_levels(cufinufft_plan plan, int levels);
int cufinufftf_set_sort_levels(cufinufftf_plan plan, int levels);
/* number of kernels this library launched in the last setpts / execute (bench.py's gpu_launches) */
int cufinufft_get_launch_counts(cufinufft_plan plan, int *out2);
int cufinufftf_get_launch_counts(cufinufftf_plan plan, int *out2);

/* ---- z-slab decomposition of ONE large 3-D transform over the GPUs of a node ----------------
 * (BASELINE.json config 5; the reference has no counterpart: its plan lives on one device,
 * src/cufinufft.cu:101-110, and its 3-D pipeline src/3d/cufinufft3d.cu:15-165 is what is split.)
 * One process per GPU; rank r of `world` owns the fine-grid planes [z0, z1) with pad = ceil(ns/2)
 * halo planes per side and the points with floor(z_rescaled) in [z0, z1) (the caller routes the
 * points: cufinufft_b200/multi.py).  The library does the device work of a rank; the two
 * exchanges of type 1 are the CALLER's (NCCL send/recv + all-reduce; multi.py uses
 * torch.distributed) so that libcufinufft.so keeps linking only cudart + cufft.
 *   slab_makeplan  like makeplan with dim = 3, ntransf = 1; cufinufft*_setpts / _destroy /
 *                  _set_stream / _set_timing / _get_timing work on the returned plan, execute does not.
 *   slab_info      out12 = {z0, z1, pad, local planes nz+2pad, nf1, nf2, nf3 (global), nf1*nf2,
 *                  rank, world, #points of the last setpts that were outside the slab (they are
 *                  pulled onto its edge: a caller error), ns}
 *   type 2 (modes -> points), NO communication: `fk` is the full mode array [mu][mt][ms]
 *                  (replicated on every rank), `c` receives the values at this rank's points.
 *   type 1 (points -> modes): slab_type1_spread(c); then for side = 0 (low halo, goes to rank-1)
 *                  and 1 (high halo, goes to rank+1): slab_halo_pack(side, sendbuf) ->
 *                  [caller sends it to the neighbour] -> on the receiver slab_halo_add(1 - side,
 *                  recvbuf); then slab_type1_finish(fk_partial): the caller SUMS fk_partial
 *                  [mu][mt][ms] over the ranks.  Halo buffers hold pad*nf1*nf2 complex numbers.
 * All pointers are device pointers; work is enqueued on the plan's stream.                     */
int cufinufft_slab_makeplan(int type, int *nmodes3, int iflag, double tol, int rank, int world,
                            cufinufft_plan *plan, cufinufft_opts *opts);
int cufinufftf_slab_makeplan(int type, int *nmodes3, int iflag, float tol, int rank, int world,
                             cufinufftf_plan *plan, cufinufft_opts *opts);
int cufinufft_slab_info(cufinufft_plan plan, long long *out12);
int cufinufftf_slab_info(cufinufftf_plan plan, long long *out12);
int cufinufft_slab_type2(cuDoubleComplex *c, cuDoubleComplex *fk, cufinufft_plan plan);
int cufinufftf_slab_type2(cuFloatComplex *c, cuFloatComplex *fk, cufinufftf_plan plan);
int cufinufft_slab_type1_spread(cuDoubleComplex *c, cufinufft_plan plan);
int cufinufftf_slab_type1_spread(cuFloatComplex *c, cufinufftf_plan plan);
int cufinufft_slab_halo_pack(int side, cuDoubleComplex *buf, cufinufft_plan plan);
int cufinufftf_slab_halo_pack(int side, cuFloatComplex *buf, cufinufftf_plan plan);
int cufinufft_slab_halo_add(int side, cuDoubleComplex *buf, cufinufft_plan plan);
int cufinufftf_slab_halo_add(int side, cuFloatComplex *buf, cufinufftf_plan plan);
int cufinufft_slab_type1_finish(cuDoubleComplex *fk_partial, cufinufft_plan plan);
int cufinufftf_slab_type1_finish(cuFloatComplex *fk_partial, cufinufftf_plan plan);

/* ---- the same decomposition with the collectives INSIDE the library (csrc/mgpu.cu): one process per GPU,
 * NCCL over NVLink.  The caller creates one communicator per process -- rank 0 calls mgpu_unique_id and hands
 * the 128 bytes to the other ranks by any means (file, MPI, a torch store) --, attaches it to its slab plan
 * and then needs three calls per transform, all stream-ordered on the plan's stream:
 *   slab_route_setpts  this rank HOLDS M points anywhere in the domain (device pointers): their owners are
 *                      computed with setpts' own rescale, counts are all-gathered, the coordinates go to the
 *                      owning ranks (grouped ncclSend/ncclRecv) and are bin-sorted there.  One host
 *                      synchronisation (buffer sizes).  slab_route_info -> {points held, points owned}.
 *   slab_route_forward / _backward   per-point complex data holder -> owner (strengths, type 1) and
 *                      owner -> holder in the holder's original order (values, type 2).
 *   slab_execute       type 2: fk (replicated, [mu][mt][ms]) -> c (owned points); no collective.
 *                      type 1: c (owned points) -> spread, ring halo exchange + add, FFTs, deconvolve,
 *                      all-reduce: fk holds the COMPLETE mode array on every rank.
 * The step-by-step calls above remain for callers that bring their own transport. */
typedef struct cufinufft_mgpu_comm_s *cufinufft_mgpu_comm;
int cufinufft_mgpu_unique_id(void *id128);
int cufinufft_mgpu_comm_create(int world, int rank, const void *id128, int device, cufinufft_mgpu_comm *comm);
int cufinufft_mgpu_comm_destroy(cufinufft_mgpu_comm comm);
int cufinufft_slab_set_comm(cufinufft_plan plan, cufinufft_mgpu_comm comm);
int cufinufftf_slab_set_comm(cufinufftf_plan plan, cufinufft_mgpu_comm comm);
int cufinufft_slab_route_setpts(int M, const double *x, const double *y, const double *z, cufinufft_plan plan);
int cufinufftf_slab_route_setpts(int M, const float *x, const float *y, const float *z, cufinufftf_plan plan);
int cufinufft_slab_route_info(cufinufft_plan plan, long long *out2);
int cufinufftf_slab_route_info(cufinufftf_plan plan, long long *out2);
int cufinufft_slab_route_forward(const cuDoubleComplex *held, cuDoubleComplex *owned, cufinufft_plan plan);
int cufinufftf_slab_route_forward(const cuFloatComplex *held, cuFloatComplex *owned, cufinufftf_plan plan);
int cufinufft_slab_route_backward(const cuDoubleComplex *owned, cuDoubleComplex *held, cufinufft_plan plan);
int cufinufftf_slab_route_backward(const cuFloatComplex *owned, cuFloatComplex *held, cufinufftf_plan plan);
int cufinufft_slab_execute(cuDoubleComplex *c, cuDoubleComplex *fk, cufinufft_plan plan);
int cufinufftf_slab_execute(cuFloatComplex *c, cuFloatComplex *fk, cufinufftf_plan plan);

#ifdef __cplusplus
}
#endif
#endif
