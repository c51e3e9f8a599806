// cfb_plan.h -- internal plan object and stage interfaces of the B200-native
// cuFINUFFT hot path.  Host C++ behind the C ABI of include/cufinufft.h.
//
// Reference counterparts: plan struct include/cufinufft_eitherprec.h:247-297,
// lifecycle src/cufinufft.cu, allocators src/memtransfer_wrapper.cu.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "../../include/cufinufft_opts.h"

namespace cfb {

constexpr int MAX_NS = 16;      // reference MAX_NSPREAD, contrib/spreadinterp.h:10
constexpr int MAX_NQUAD = 100;  // reference contrib/common.h:10

template <typename T> struct cplx_of;
template <> struct cplx_of<float>  { using type = float2; };
template <> struct cplx_of<double> { using type = double2; };

// One owned device allocation that only ever grows (setpts may be called repeatedly
// on a plan; the reference frees and re-mallocs every time, memtransfer_wrapper.cu:238-279).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename U> U *as() const { return static_cast<U *>(p); }
};

// One sorted point: rescaled coordinates (fine-grid units) + the point's index in the caller's
// arrays.  16 bytes (fp32) / 32 bytes (fp64): one vector load per point in spread / interp and
// one sector-sized scattered store per point in setpts.
template <typename T> struct PtRec;
template <> struct alignas(16) PtRec<float>  { float x, y, z; int idx; };
template <> struct alignas(32) PtRec<double> { double x, y, z; long long idx; };

// geometry of the setpts sort key: bin grid and the stencil-cell grid inside a bin
struct SortGeo {
    int nf[3], bs[3], nb[3];
    // internal bins (work items of the tile engines): ibs divides bs, spb = bs / ibs sub-bins per
    // reference bin and dimension; internal bin = refbin * spbt + sub, so the reference's bin-major
    // order is kept and its arrays are sums over groups of spbt internal bins
    int ibs[3], spb[3], spbt;
    // power-of-two bin sizes (the defaults: 32x32, 16x16x2): lg_ibs >= 0, and every division of the key is an
    // exact scaling by inv_ibs = 1 / ibs -- same integers as the divisions, a third of the instructions
    int lg_ibs[3], lg_spb[3];
    float inv_ibs[3];
    // stencil-origin cells per internal bin and dimension (the "fine" part of the order).  Either
    // they are part of the global key (nk = nkf, cpb = cpbf) or the global key is the bin alone
    // (nk = 1, cpb = 1) and every work item is sorted by cell afterwards (local_sort_kernel)
    int nkf[3], cpbf;
    int nk[3];      // distinct stencil origins per bin and dimension (1 = key is the bin alone)
    int cpb;        // nk[0]*nk[1]*nk[2]
    int ns;
    // z-slab plans (slab.cu): z is rescaled on the GLOBAL fine grid nfz and binned in slab-local
    // planes zl = z_r - zshift; nf[2] is then the local plane count.  Ordinary plans: nfz = nf[2],
    // zshift = 0, zlo = 0, zhi = nf[2].
    int nfz, zshift, zlo, zhi;
    // bank-class order (type-2 plans served by the tile interpolation engine): the key inside a bin is the
    // point's shared-memory bank class, (its first stencil cell's index in the bin tile) mod bankc, instead of
    // the stencil cell; the tile has bex x bey (x bez) cells and a halo of bpad.  bankc = 0: stencil-cell order.
    int bankc, bex, bey, bez, bpad;
};

// multi-GPU state of a slab plan (mgpu.cu): communicator + routing of the points between the rank that
// holds them and the rank that owns them
struct MgpuComm;
struct RouteState {
    MgpuComm *comm = nullptr;
    int n_held = 0, n_owned = 0;
    DevBuf slot;                         // int[n_held]: position of held point i in the send order (grouped by owner)
    DevBuf owned[3];                     // T[n_owned]: coordinates of the owned points (what setpts sorts)
    DevBuf sendbuf, recvbuf;             // staging for one exchanged array
    DevBuf counts;                       // int[world*world + 2*world] device scratch
    DevBuf halo[4];                      // send_lo, send_hi, recv_prev, recv_next
    long long send_counts[64] = {0}, recv_counts[64] = {0};
};

template <typename T>
struct Plan {
    using C = typename cplx_of<T>::type;

    cufinufft_opts opts;
    int type = 0, dim = 0;
    int ms = 1, mt = 1, mu = 1;          // modes
    int nf1 = 1, nf2 = 1, nf3 = 1;       // fine grid
    int ntransf = 1, maxbatch = 1, iflag = 1;
    // kernel
    int ns = 0;
    T es_beta = 0, es_c = 0, es_halfwidth = 0;
    // bins
    int bs[3] = {1, 1, 1};
    int nbin[3] = {1, 1, 1};
    int nbins = 1;
    int method = 0;                      // effective engine: 1 = GM/GM-sort, 2 = SM tiles
    bool sorted = true;
    // internal bins: the tile engines' work items may be finer than the reference's bins (smaller
    // per-warp tiles -> more resident warps); chosen in setpts, ibs == bs when not worth it
    int ibs[3] = {1, 1, 1};
    int spb[3] = {1, 1, 1};
    int nibins = 1;
    int imaxsub = 1024;                  // points per work item of the internal list (>= gpu_maxsubprobsize)
    bool ilist = false;                  // the engines use their own work list (isubstart / is2b) instead of the reference's
    DevBuf isubstart, is2b;              // internal subproblem offsets [nibins+1] and map (unused when ibs == bs)
    // points (borrowed) + derived (owned)
    int M = -1;
    const T *kx = nullptr, *ky = nullptr, *kz = nullptr;
    DevBuf recs;                         // PtRec<T>[M], sorted by (bin, stencil cell)
    DevBuf idxnupts;                     // int[M]: inverse permutation (materialised on demand)
    DevBuf keyoff, tilesum;              // int[4 + nkeys + 1] key histogram -> offsets -> cursors (setpts.cu); scan scratch
    const int *key_offsets() const { return keyoff.as<int>() + 3; }   // [k] = first sorted point of key k, [nkeys] = M (after setpts)
    SortGeo sortgeo;
    bool fine_sort_allowed = true;
    bool local_sort = false;             // two-level order: bins globally, stencil cells per work item
    int sort_levels = 0;                 // 0 automatic, 1 / 2 forced (cufinufft*_set_sort_levels: tests, A/B)
    int sort_partition = 0;              // coarse partition in front of the counting sort: 0 automatic, 1 never, 2 always
    long long sort_bucket_bytes = 4LL << 20;    // records per coarse bucket (bytes): a window L2 holds several of
    bool key_generic = false;            // force the generic key code (tests: cufinufft*_set_sort_levels + 16)
    bool partitioned = false;            // ... used by the last setpts
    DevBuf tmprecs, coarse;              // PtRec<T>[M] in coarse-bucket order; int[<= 4097] bucket counts / cursors
    bool idx_valid = false;
    DevBuf binsize, binstartpts, numsubprob, subprobstartpts, subprob_to_bin;
    DevBuf scalars;                      // int[8]: [0] totalnumsubprob, [1] work counter, ...
    DevBuf fw;                           // C[maxbatch * nf1*nf2*nf3]
    DevBuf fwker[3];                     // T[nf/2+1]
    DevBuf hostside;                     // device staging for the *_host convenience calls
    DevBuf hcoef;                        // Horner coefficients [ncoef][ns] as T (kerevalmeth=1)
    int horner_ncoef = 0;
    cufftHandle fftplan = 0;
    bool have_fft = false;
    cudaStream_t stream = cudaStreamPerThread;   // the reference is built with --default-stream per-thread (Makefile:35)
    int device = 0;
    int num_sms = 148;
    int max_smem_optin = 227 * 1024;
    long long l2_bytes = 126LL << 20;
    // SM-tile geometry (set at makeplan)
    int tile_pad = 0;                    // ceil(ns/2)
    int tile_sy = 0, tile_sz = 0;        // padded strides (cells)
    int tile_cells = 0;
    int tile_cost = 0;                   // shared-memory wavefronts per point of the chosen tile layout (bank-conflict model)
    int sm2_rmc = 0;                     // spread_sm2_kernel<3, 6>: row-order class chosen with the tile strides (spread_sm2.cuh: sm2_rowmap_tab)
    int sm_warps = 0;                    // warps per block for the SM spread kernel (0 = SM unusable)
    bool plane_engine = false;           // spread_plane.cuh serves this plan's points (chosen in setpts)
    int interp_engine = 0;               // 0 auto (tile when sorted and it fits), 1 gather, 2 tile
    int bank_classes = 0;                // > 0: the points of the last setpts are in bank-class order (setpts.cu)
    // z-slab decomposition of one 3-D transform (slab.cu): this plan owns the fine-grid
    // planes [z0, z1) of a global grid with nf3g planes and holds them with `tile_pad` halo planes on
    // both sides: nf3 = z1 - z0 + 2*tile_pad local planes, local plane l = global plane zshift + l.
    bool slab = false;
    int nf3g = 0, z0 = 0, z1 = 0, zshift = 0;
    int slab_rank = 0, slab_world = 1;
    cufftHandle fft2d = 0, fftz = 0;     // batched (x,y) planes of the local grid; strided z pencils of zbuf
    bool have_fft2d = false, have_fftz = false;
    DevBuf zbuf;                         // C[nf3g][mt][ms]: mode columns, all z planes (z-pencil layout)
    RouteState route;                    // multi-GPU (mgpu.cu)
    // host-pointer calls (plan.cu: setpts_host / execute_host): the points are ALSO sorted in `chunks` of the
    // caller's index range (child plans without grid or FFT), so that the H2D copy of the strengths of chunk
    // k+1 overlaps the spreading of chunk k (type 1) and the D2H copy of the values of chunk k the
    // interpolation of chunk k+1 (type 2)
    double tol = 0;
    std::vector<Plan<T> *> chunks;
    std::vector<long long> chunk_off;    // first point of every chunk, + the total
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> chunk_ev;
    // timing / accounting
    bool timing = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int launches_setpts = 0, launches_exec = 0;

    size_t grid_cells() const { return (size_t)nf1 * nf2 * nf3; }
    int nf3_global() const { return slab ? nf3g : nf3; }
    size_t nmodes() const { return (size_t)ms * mt * mu; }
};

// ---- host math (hostmath.cpp) --------------------------------------------
template <typename T>
int setup_spreader(T eps, double upsampfac, int kerevalmeth, int *ns, T *beta, T *halfwidth, T *c);
int next235beven(int n, int b);
int set_nf_type12(int m, double upsampfac, int ns, int gpu_method, int obinsize);
void gauss_legendre(int n, double *x, double *w);
template <typename T>
void fseries_precomp(int nf, int ns, T beta, T es_c, T halfwidth, T *f, double *a_reim);

// ---- device stages ----------------------------------------------------------
template <typename T> int stage_fseries(Plan<T> &p);                       // deconv.cu
template <typename T> int stage_setpts(Plan<T> &p);                        // setpts.cu
template <typename T> int materialize_idxnupts(Plan<T> &p);                // setpts.cu
template <typename T> int stage_spread(Plan<T> &p, const typename Plan<T>::C *c, typename Plan<T>::C *fw, int nt);
template <typename T> int stage_interp(Plan<T> &p, typename Plan<T>::C *c, const typename Plan<T>::C *fw, int nt);
template <typename T> int stage_deconvolve(Plan<T> &p, typename Plan<T>::C *fk, const typename Plan<T>::C *fw, int nt);
template <typename T> int stage_amplify(Plan<T> &p, const typename Plan<T>::C *fk, typename Plan<T>::C *fw, int nt);
template <typename T> void plan_tile_geometry(Plan<T> &p);                 // spread.cu
template <typename T> void choose_internal_bins(Plan<T> &p, long long M);  // spread.cu
template <typename T> bool interp_tile_applies(const Plan<T> &p);          // spread.cu: the tile interpolation engine will serve this plan
// z-slab stages (slab.cu)
template <typename T> int slab_make_ffts(Plan<T> &p);
template <typename T> int slab_type2(Plan<T> &p, typename Plan<T>::C *c, const typename Plan<T>::C *fk);
template <typename T> int slab_type1_spread(Plan<T> &p, const typename Plan<T>::C *c);
template <typename T> int slab_halo_pack(Plan<T> &p, int side, typename Plan<T>::C *buf);
template <typename T> int slab_halo_add(Plan<T> &p, int side, const typename Plan<T>::C *buf);
template <typename T> int slab_type1_finish(Plan<T> &p, typename Plan<T>::C *fk_partial);

#define CFB_CUDA_OK(call)                                                              \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            fprintf(stderr, "[cufinufft-b200] CUDA error %s at %s:%d: %s\n", #call,    \
                    __FILE__, __LINE__, cudaGetErrorString(e__));                      \
            return 11;                                                                 \
        }                                                                              \
    } while (0)

}  // namespace cfb

// the opaque handles of the C ABI
struct cufinufft_plan_s  { cfb::Plan<double> *p; cfb::DevBuf dc, dfk; };
struct cufinufftf_plan_s { cfb::Plan<float> *p;  cfb::DevBuf dc, dfk; };
