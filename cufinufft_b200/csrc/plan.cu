// plan.cu -- plan lifecycle and the extern "C" ABI of libcufinufft.so.
//
// Reference: src/cufinufft.cu (makeplan :78-273, setpts :275-493, execute :495-569,
// destroy :571-637, default_opts :639-730), the batch loops src/{1,2,3}d/cufinufft*d.cu,
// allocation src/memtransfer_wrapper.cu.  Host code only; every device stage is in
// setpts.cu / spread*.cu / deconv.cu.  Differences from the reference that a caller can
// observe are listed in include/cufinufft.h.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include "cfb_plan.h"
#include "../../include/cufinufft.h"
#include "../../include/cufinufft_b200.h"
#include "horner_coeffs.inc"

namespace cfb {

struct DeviceGuard {            // every API call runs on opts.gpu_device_id and restores the
    int prev = 0;               // caller's device (reference src/cufinufft.cu:101-110,270)
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (dev != prev) cudaSetDevice(dev);
        target = dev;
        // A stale non-sticky error left by other CUDA code in the process would be reported by our own
        // post-launch checks as ours: take it off the error slot, but say so instead of hiding it.
        const cudaError_t stale = cudaGetLastError();
        if (stale != cudaSuccess)
            fprintf(stderr, "[cufinufft-b200] note: a CUDA error was already pending on entry (not ours): %s\n", cudaGetErrorString(stale));
    }
    ~DeviceGuard() { if (target != prev) cudaSetDevice(prev); }
    int target = 0;
};

static int default_opts(int type, int dim, cufinufft_opts *o)
{
    if (!o) return CFB_ERR_BAD_ARG;
    o->upsampfac = 2.0;
    o->gpu_nstreams = 0;
    o->gpu_sort = 1;
    o->gpu_maxsubprobsize = 1024;
    o->gpu_obinsizex = o->gpu_obinsizey = o->gpu_obinsizez = -1;
    o->gpu_binsizex = o->gpu_binsizey = o->gpu_binsizez = -1;
    o->gpu_spreadinterponly = 0;
    o->gpu_kerevalmeth = 0;
    o->gpu_method = 0;
    o->gpu_device_id = 0;
    if (dim < 1 || dim > 3) return CFB_ERR_BAD_ARG;
    if (type == 1) o->gpu_method = 2;
    else if (type == 2) o->gpu_method = 1;
    else { fprintf(stderr, "[cufinufft-b200] type %d: Not Implemented yet\n", type); return 1; }
    return 0;
}

// SETUP_BINSIZE, src/cufinufft.cu:17-73
static void setup_binsize(int dim, cufinufft_opts *o)
{
    if (dim == 1) {
        if (o->gpu_binsizex < 0) o->gpu_binsizex = 1024;
        o->gpu_binsizey = 1; o->gpu_binsizez = 1;
    } else if (dim == 2) {
        if (o->gpu_binsizex < 0) o->gpu_binsizex = 32;
        if (o->gpu_binsizey < 0) o->gpu_binsizey = 32;
        o->gpu_binsizez = 1;
    } else if (o->gpu_method == 4) {
        if (o->gpu_obinsizex < 0) o->gpu_obinsizex = 8;
        if (o->gpu_obinsizey < 0) o->gpu_obinsizey = 8;
        if (o->gpu_obinsizez < 0) o->gpu_obinsizez = 8;
        if (o->gpu_binsizex < 0) o->gpu_binsizex = 4;
        if (o->gpu_binsizey < 0) o->gpu_binsizey = 4;
        if (o->gpu_binsizez < 0) o->gpu_binsizez = 4;
    } else {
        if (o->gpu_binsizex < 0) o->gpu_binsizex = 16;
        if (o->gpu_binsizey < 0) o->gpu_binsizey = 16;
        if (o->gpu_binsizez < 0) o->gpu_binsizez = 2;
    }
}

template <typename T>
static void free_plan(Plan<T> *p)
{
    if (!p) return;
    for (Plan<T> *c : p->chunks) free_plan(c);
    p->chunks.clear();
    for (cudaEvent_t e : p->chunk_ev) cudaEventDestroy(e);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->have_fft) cufftDestroy(p->fftplan);
    if (p->have_fft2d) cufftDestroy(p->fft2d);
    if (p->have_fftz) cufftDestroy(p->fftz);
    p->zbuf.release();
    for (DevBuf *b : {&p->recs, &p->keyoff, &p->tilesum, &p->idxnupts, &p->binsize, &p->binstartpts, &p->numsubprob,
                      &p->subprobstartpts, &p->subprob_to_bin, &p->isubstart, &p->is2b, &p->scalars, &p->fw, &p->fwker[0], &p->fwker[1],
                      &p->fwker[2], &p->hostside, &p->hcoef, &p->tmprecs, &p->coarse, &p->route.slot, &p->route.owned[0], &p->route.owned[1], &p->route.owned[2],
                      &p->route.sendbuf, &p->route.recvbuf, &p->route.counts, &p->route.halo[0], &p->route.halo[1], &p->route.halo[2],
                      &p->route.halo[3]})
        b->release();
    for (auto &e : p->ev) if (e) cudaEventDestroy(e);
    delete p;
}

// Host-only part of makeplan (no CUDA call): validates the arguments and derives every
// plan-time number -- kernel width/beta, fine grid, bin grid, batch size, engine choice.
// Reference: src/cufinufft.cu:118-175 (+ SETUP_BINSIZE :17-73, spread wrappers' numbins).
template <typename T>
static int plan_host_setup(Plan<T> *p, int type, int dim, const int *nmodes, int iflag, int ntransf, T tol,
                           int maxbatchsize, const cufinufft_opts *user_opts, int slab_rank = -1, int slab_world = 0)
{
    if (!nmodes || dim < 1 || dim > 3 || ntransf < 1) return CFB_ERR_BAD_ARG;
    if (type == 3) { fprintf(stderr, "[cufinufft-b200] type 3: Not Implemented yet\n"); return CFB_ERR_NOT_IMPLEMENTED; }
    if (type != 1 && type != 2) return CFB_ERR_BAD_ARG;
    if (user_opts) p->opts = *user_opts;
    else default_opts(type, dim, &p->opts);
    p->device = p->opts.gpu_device_id;

    p->tol = (double)tol;
    int ier = setup_spreader<T>(tol, p->opts.upsampfac, p->opts.gpu_kerevalmeth, &p->ns, &p->es_beta, &p->es_halfwidth, &p->es_c);
    if (ier > 1) return ier;

    p->type = type; p->dim = dim;
    p->ms = nmodes[0];
    p->mt = dim > 1 ? nmodes[1] : 1;
    p->mu = dim > 2 ? nmodes[2] : 1;
    if (p->ms < 1 || p->mt < 1 || p->mu < 1) return CFB_ERR_BAD_ARG;
    int m = p->opts.gpu_method;
    if (m < 1 || m > 4 || (m == 3 && dim != 2) || (m == 4 && dim != 3)) {
        fprintf(stderr, "[cufinufft-b200] invalid gpu_method %d for dim %d\n", m, dim);
        return CFB_ERR_BAD_ARG;
    }
    setup_binsize(dim, &p->opts);
    p->nf1 = set_nf_type12(p->ms, p->opts.upsampfac, p->ns, m, p->opts.gpu_obinsizex);
    if (dim > 1) p->nf2 = set_nf_type12(p->mt, p->opts.upsampfac, p->ns, m, p->opts.gpu_obinsizey);
    if (dim > 2) p->nf3 = set_nf_type12(p->mu, p->opts.upsampfac, p->ns, m, p->opts.gpu_obinsizez);
    if (slab_world > 0) {
        // z-slab share of one 3-D transform (slab.cu): planes [z0, z1) of the global grid + pad halo
        // planes per side; every slab must be at least one halo thick (single-neighbour exchange)
        if (dim != 3 || ntransf != 1 || slab_rank < 0 || slab_rank >= slab_world) return CFB_ERR_BAD_ARG;
        const int pad = (p->ns + 1) / 2;
        const int base = p->nf3 / slab_world, extra = p->nf3 % slab_world;
        if (base < pad) { fprintf(stderr, "[cufinufft-b200] slab thinner than the kernel halo\n"); return CFB_ERR_BAD_ARG; }
        p->slab = true; p->slab_rank = slab_rank; p->slab_world = slab_world;
        p->nf3g = p->nf3;
        p->z0 = slab_rank * base + (slab_rank < extra ? slab_rank : extra);
        p->z1 = p->z0 + base + (slab_rank < extra ? 1 : 0);
        p->zshift = p->z0 - pad;
        p->nf3 = p->z1 - p->z0 + 2 * pad;                                          // local planes
        if ((double)p->nf3g * p->ms * p->mt > 2147483647.0) return CFB_ERR_BAD_ARG;
    }
    if ((double)p->nf1 * p->nf2 * p->nf3 > 2147483647.0) return CFB_ERR_BAD_ARG;   // int32 cells, as the reference
    p->iflag = iflag >= 0 ? 1 : -1;
    p->ntransf = ntransf;
    if (maxbatchsize <= 0) maxbatchsize = ntransf < 8 ? ntransf : 8;     // reference heuristic, :167-168
    p->maxbatch = maxbatchsize;

    // Method 4 (3-D block gather) and 3 (2-D "Paul") are alternative *schedules* of the same
    // spread in the reference; here their requests are served by the SM tile engine on the
    // method's own fine-grid size and bin size (scope).
    p->method = (m == 1) ? 1 : 2;
    p->sorted = (m != 1) || (p->opts.gpu_sort != 0);
    p->bs[0] = p->opts.gpu_binsizex; p->bs[1] = dim > 1 ? p->opts.gpu_binsizey : 1; p->bs[2] = dim > 2 ? p->opts.gpu_binsizez : 1;
    if (m == 4) { p->bs[0] = p->opts.gpu_obinsizex; p->bs[1] = p->opts.gpu_obinsizey; p->bs[2] = p->opts.gpu_obinsizez; }
    for (int d = 0; d < dim; ++d)
        if (p->bs[d] < 1) { fprintf(stderr, "[cufinufft-b200] invalid bin size\n"); return CFB_ERR_BAD_ARG; }
    if (p->opts.gpu_maxsubprobsize < 1) return CFB_ERR_BAD_ARG;
    const int nf[3] = {p->nf1, p->nf2, p->nf3};
    p->nbins = 1;
    for (int d = 0; d < 3; ++d) {
        p->nbin[d] = d < dim ? (int)ceil((T)nf[d] / p->bs[d]) : 1;       // numbins = ceil((FLT)nf/bin), spread2d_wrapper.cu:405-406
        p->nbins *= p->nbin[d];
        p->ibs[d] = p->bs[d]; p->spb[d] = 1;
    }
    p->nibins = p->nbins;
    return 0;
}

template <typename T>
static int makeplan(int type, int dim, int *nmodes, int iflag, int ntransf, T tol, int maxbatchsize,
                    Plan<T> **out, cufinufft_opts *user_opts, int slab_rank = -1, int slab_world = 0)
{
    if (!out) return CFB_ERR_BAD_ARG;
    *out = nullptr;
    Plan<T> *p = new (std::nothrow) Plan<T>();
    if (!p) return CFB_ERR_BAD_ARG;
    if (int ier = plan_host_setup<T>(p, type, dim, nmodes, iflag, ntransf, tol, maxbatchsize, user_opts, slab_rank, slab_world)) { delete p; return ier; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        fprintf(stderr, "[cufinufft-b200] no CUDA device available: this library has no CPU fallback\n");
        delete p;
        return CFB_ERR_CUDA;
    }
    if (p->device < 0 || p->device >= ndev) { delete p; return CFB_ERR_BAD_ARG; }
    DeviceGuard guard(p->device);
    const int nf[3] = {p->nf1, p->nf2, p->nf3_global()};

    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, p->device) != cudaSuccess) { delete p; return CFB_ERR_CUDA; }
    p->num_sms = prop.multiProcessorCount;
    p->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    p->l2_bytes = prop.l2CacheSize;
    plan_tile_geometry(*p);

    auto fail = [&](int code) { free_plan(p); return code; };
    const size_t nb = (size_t)p->nbins;
    if (p->binsize.reserve(nb * sizeof(int)) || p->binstartpts.reserve(nb * sizeof(int)) ||
        p->numsubprob.reserve(nb * sizeof(int)) || p->subprobstartpts.reserve((nb + 1) * sizeof(int)) ||
        p->scalars.reserve(8 * sizeof(int)))
        return fail(CFB_ERR_CUDA);
    cudaMemset(p->scalars.p, 0, 8 * sizeof(int));
    if (p->opts.gpu_kerevalmeth == 1) {
        p->horner_ncoef = cfb_horner_ncoef[p->ns];
        T tab[18 * 16];
        for (int k = 0; k < p->horner_ncoef; ++k)
            for (int i = 0; i < p->ns; ++i) tab[k * p->ns + i] = (T)cfb_horner_coeffs[p->ns][k * 16 + i];
        if (p->hcoef.reserve(sizeof(tab))) return fail(CFB_ERR_CUDA);
        if (cudaMemcpy(p->hcoef.p, tab, sizeof(T) * p->horner_ncoef * p->ns, cudaMemcpyHostToDevice)) return fail(CFB_ERR_CUDA);
    }
    if (!p->opts.gpu_spreadinterponly) {
        if (p->fw.reserve((size_t)p->maxbatch * p->grid_cells() * sizeof(typename Plan<T>::C))) return fail(CFB_ERR_CUDA);
        for (int d = 0; d < dim; ++d)
            if (p->fwker[d].reserve((size_t)(nf[d] / 2 + 1) * sizeof(T))) return fail(CFB_ERR_CUDA);
        if (p->slab) {
            if (p->zbuf.reserve((size_t)p->nf3g * p->mt * p->ms * sizeof(typename Plan<T>::C))) return fail(CFB_ERR_CUDA);
            if (int e = slab_make_ffts(*p)) { fprintf(stderr, "[cufinufft-b200] slab cufftPlanMany failed\n"); return fail(e); }
        } else {
            int n[3];
            for (int d = 0; d < dim; ++d) n[d] = nf[dim - 1 - d];            // slowest first: {nf3, nf2, nf1}
            int dist = (int)p->grid_cells();
            cufftResult fr = cufftPlanMany(&p->fftplan, dim, n, n, 1, dist, n, 1, dist,
                                           sizeof(T) == 4 ? CUFFT_C2C : CUFFT_Z2Z, p->maxbatch);
            if (fr != CUFFT_SUCCESS) { fprintf(stderr, "[cufinufft-b200] cufftPlanMany failed (%d)\n", (int)fr); return fail(CFB_ERR_CUFFT); }
            p->have_fft = true;
        }
        if (stage_fseries(*p)) return fail(CFB_ERR_CUDA);
    }
    *out = p;
    return 0;            // the eps warning (1) is swallowed exactly as the reference does
}

template <typename T>
static int setpts(Plan<T> *p, int M, const T *x, const T *y, const T *z)
{
    if (!p || M < 0) return CFB_ERR_BAD_ARG;
    if (M > 0 && (!x || (p->dim > 1 && !y) || (p->dim > 2 && !z))) return CFB_ERR_BAD_ARG;
    DeviceGuard guard(p->device);
    p->M = M;
    p->kx = x; p->ky = p->dim > 1 ? y : nullptr; p->kz = p->dim > 2 ? z : nullptr;
    return stage_setpts(*p);
}

static cufftResult fft_exec(cufftHandle h, float2 *d, int dir) { return cufftExecC2C(h, d, d, dir); }
static cufftResult fft_exec(cufftHandle h, double2 *d, int dir) { return cufftExecZ2Z(h, d, d, dir); }

template <typename T>
static int execute(Plan<T> *p, typename Plan<T>::C *c, typename Plan<T>::C *fk)
{
    using C = typename Plan<T>::C;
    if (!p) return CFB_ERR_BAD_ARG;
    if (p->M < 0) { fprintf(stderr, "[cufinufft-b200] execute before setpts\n"); return CFB_ERR_NO_POINTS_SET; }
    if (p->slab) { fprintf(stderr, "[cufinufft-b200] slab plans execute through cufinufft*_slab_* (include/cufinufft_b200.h)\n"); return CFB_ERR_BAD_ARG; }
    DeviceGuard guard(p->device);
    cudaStream_t st = p->stream;
    p->launches_exec = 0;
    const size_t cells = p->grid_cells();
    auto mark = [&](int i) { if (p->timing) cudaEventRecord(p->ev[i], st); };
    if (p->timing) for (auto &e : p->ev) if (!e) cudaEventCreate(&e);

    if (p->opts.gpu_spreadinterponly) {        // fk IS the fine grid [ntransf][nf3][nf2][nf1]
        for (int b0 = 0; b0 < p->ntransf; b0 += p->maxbatch) {
            int nt = std::min(p->maxbatch, p->ntransf - b0);
            C *cb = c + (size_t)b0 * p->M, *fb = fk + (size_t)b0 * cells;
            if (p->type == 1) {
                CFB_CUDA_OK(cudaMemsetAsync(fb, 0, (size_t)nt * cells * sizeof(C), st));
                if (int e = stage_spread(*p, cb, fb, nt)) return e;
            } else if (int e = stage_interp(*p, cb, fb, nt)) return e;
        }
        return 0;
    }
    CFB_CUDA_OK((cudaError_t)(cufftSetStream(p->fftplan, st) == CUFFT_SUCCESS ? cudaSuccess : cudaErrorUnknown));
    C *fw = p->fw.template as<C>();
    for (int b0 = 0; b0 < p->ntransf; b0 += p->maxbatch) {      // batch loop, src/2d/cufinufft2d.cu:39-90
        int nt = std::min(p->maxbatch, p->ntransf - b0);
        C *cb = c + (size_t)b0 * p->M, *fb = fk + (size_t)b0 * p->nmodes();
        if (p->type == 1) {
            mark(0);
            CFB_CUDA_OK(cudaMemsetAsync(fw, 0, (size_t)p->maxbatch * cells * sizeof(C), st));
            mark(1);
            if (int e = stage_spread(*p, cb, fw, nt)) return e;
            mark(2);
            if (fft_exec(p->fftplan, fw, p->iflag) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
            mark(3);
            if (int e = stage_deconvolve(*p, fb, fw, nt)) return e;
            mark(4);
        } else {
            mark(0);
            if (nt < p->maxbatch)               // short last batch: cuFFT still transforms maxbatch grids
                CFB_CUDA_OK(cudaMemsetAsync(fw + (size_t)nt * cells, 0, (size_t)(p->maxbatch - nt) * cells * sizeof(C), st));
            mark(1);
            if (int e = stage_amplify(*p, fb, fw, nt)) return e;
            mark(2);
            if (fft_exec(p->fftplan, fw, p->iflag) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
            mark(3);
            if (int e = stage_interp(*p, cb, fw, nt)) return e;
            mark(4);
        }
    }                                            // launches_exec counts OUR kernels only (cuFFT's are not)
    return 0;
}

template <typename T>
static int destroy(Plan<T> *p)
{
    if (!p) return 1;
    DeviceGuard guard(p->device);
    cudaStreamSynchronize(p->stream);
    free_plan(p);
    return 0;
}

// ---- extensions --------------------------------------------------------------
// Host-pointer calls.  Besides the plan's own sort (so that every other call keeps working), large single
// transforms get their points sorted a second time in K chunks of the CALLER's index range -- child plans
// that own only sort state (gpu_spreadinterponly: no grid, no FFT) and spread into / interpolate from the
// parent's grid.  execute_host then overlaps the PCIe copies of the strengths / values with the kernels chunk by
// chunk; with one chunk (small inputs, batches, slab plans) it is copy -> execute -> copy.
// (VERDICT r1 "weak 15": the host-buffer step was 75 % PCIe, nothing overlapped.)
constexpr int PIPE_MIN_POINTS = 2000000;

template <typename T>
static int setpts_host(Plan<T> *p, int M, const T *x, const T *y, const T *z)
{
    if (!p || M < 0) return CFB_ERR_BAD_ARG;
    if (M > 0 && (!x || (p->dim > 1 && !y) || (p->dim > 2 && !z))) return CFB_ERR_BAD_ARG;
    DeviceGuard guard(p->device);
    const size_t n = (size_t)(M > 0 ? M : 1);
    CFB_CUDA_OK(p->hostside.reserve(n * sizeof(T) * p->dim));
    T *d = p->hostside.template as<T>();
    const T *src[3] = {x, y, z};
    for (int k = 0; k < p->dim; ++k)
        if (M > 0) CFB_CUDA_OK(cudaMemcpyAsync(d + (size_t)k * n, src[k], (size_t)M * sizeof(T), cudaMemcpyHostToDevice, p->stream));
    if (int e = setpts(p, M, d, d + n, d + 2 * n)) return e;

    // Type 1 pays for chunks: the spread kernels accumulate runs of points that share a stencil, and dealing the
    // points of a cell to K chunks divides the run length by K.  2-D: 4 chunks still win (config 1: 2.11 -> 1.85 ms
    // per host-buffer step); 3-D, where a run saves ns^3 tile updates: two chunks (config 3, clustered: 25.9 -> 24.0 ms;
    // three give 29.1, four 31.4; profiles/r02x, r03q).  Type 2 has no such coupling (config 2: 16.9 -> 14.2 ms).
    static const int pipe3d = [] { const char *e = getenv("CFB_PIPE3D"); return e && atoi(e) > 0 ? atoi(e) : 2; }();   // experiments
    const int K = (M >= PIPE_MIN_POINTS && p->ntransf == 1 && !p->slab && !p->opts.gpu_spreadinterponly)
                      ? (p->type == 1 ? (p->dim <= 2 ? 4 : pipe3d) : 8) : 1;
    if (K == 1) {
        for (Plan<T> *c : p->chunks) free_plan(c);
        p->chunks.clear();
        p->chunk_off.clear();
        return 0;
    }
    if ((int)p->chunks.size() != K) {
        for (Plan<T> *c : p->chunks) free_plan(c);
        p->chunks.clear();
        int nm[3] = {p->ms, p->mt, p->mu};
        cufinufft_opts o = p->opts;
        o.gpu_spreadinterponly = 1;
        for (int k = 0; k < K; ++k) {
            Plan<T> *c = nullptr;
            if (int e = makeplan<T>(p->type, p->dim, nm, p->iflag, 1, (T)p->tol, 1, &c, &o)) return e;
            p->chunks.push_back(c);
        }
        if (!p->copy_stream) CFB_CUDA_OK(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
        while ((int)p->chunk_ev.size() < K) {
            cudaEvent_t ev;
            CFB_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            p->chunk_ev.push_back(ev);
        }
    }
    p->chunk_off.assign(K + 1, 0);
    for (int k = 0; k <= K; ++k) p->chunk_off[k] = (long long)M * k / K;
    for (int k = 0; k < K; ++k) {
        Plan<T> *c = p->chunks[k];
        c->stream = p->stream;
        c->interp_engine = p->interp_engine;
        const long long o0 = p->chunk_off[k];
        if (int e = setpts(c, (int)(p->chunk_off[k + 1] - o0), d + o0, d + n + o0, d + 2 * n + o0)) return e;
    }
    return 0;
}

template <typename T>
static int execute_host(Plan<T> *p, typename Plan<T>::C *c, typename Plan<T>::C *fk, DevBuf &dc, DevBuf &dfk)
{
    using C = typename Plan<T>::C;
    if (!p || p->M < 0) return CFB_ERR_BAD_ARG;
    DeviceGuard guard(p->device);
    const size_t nc = (size_t)p->ntransf * (size_t)p->M * sizeof(C);      // M == 0: nothing of c is read or written
    const size_t nk = (size_t)p->ntransf * (p->opts.gpu_spreadinterponly ? p->grid_cells() : p->nmodes()) * sizeof(C);
    if ((nc && !c) || !fk) return CFB_ERR_BAD_ARG;
    CFB_CUDA_OK(dc.reserve(nc ? nc : sizeof(C)));
    CFB_CUDA_OK(dfk.reserve(nk));
    cudaStream_t st = p->stream;
    const int K = (int)p->chunks.size();
    if (K > 1 && (long long)p->M == p->chunk_off[K]) {
        // ---- chunked pipeline (one transform): copies on copy_stream, kernels on the plan's stream
        cudaStream_t cs = p->copy_stream;
        C *fw = p->fw.template as<C>(), *dcp = dc.as<C>(), *dkp = dfk.as<C>();
        const size_t cells = p->grid_cells();
        CFB_CUDA_OK((cudaError_t)(cufftSetStream(p->fftplan, st) == CUFFT_SUCCESS ? cudaSuccess : cudaErrorUnknown));
        p->launches_exec = 0;
        if (p->type == 1) {
            for (int k = 0; k < K; ++k) {
                const long long o0 = p->chunk_off[k], m = p->chunk_off[k + 1] - o0;
                CFB_CUDA_OK(cudaMemcpyAsync(dcp + o0, c + o0, (size_t)m * sizeof(C), cudaMemcpyHostToDevice, cs));
                CFB_CUDA_OK(cudaEventRecord(p->chunk_ev[k], cs));
            }
            CFB_CUDA_OK(cudaMemsetAsync(fw, 0, (size_t)p->maxbatch * cells * sizeof(C), st));
            for (int k = 0; k < K; ++k) {
                Plan<T> *ch = p->chunks[k];
                ch->stream = st;
                CFB_CUDA_OK(cudaStreamWaitEvent(st, p->chunk_ev[k], 0));
                if (int e = stage_spread(*ch, dcp + p->chunk_off[k], fw, 1)) return e;
                p->launches_exec += ch->launches_exec; ch->launches_exec = 0;
            }
            if (fft_exec(p->fftplan, fw, p->iflag) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
            if (int e = stage_deconvolve(*p, dkp, fw, 1)) return e;
            CFB_CUDA_OK(cudaMemcpyAsync(fk, dkp, nk, cudaMemcpyDeviceToHost, st));
        } else {
            CFB_CUDA_OK(cudaMemcpyAsync(dkp, fk, nk, cudaMemcpyHostToDevice, st));
            if (p->maxbatch > 1) CFB_CUDA_OK(cudaMemsetAsync(fw + cells, 0, (size_t)(p->maxbatch - 1) * cells * sizeof(C), st));
            if (int e = stage_amplify(*p, dkp, fw, 1)) return e;
            if (fft_exec(p->fftplan, fw, p->iflag) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
            for (int k = 0; k < K; ++k) {
                Plan<T> *ch = p->chunks[k];
                ch->stream = st;
                const long long o0 = p->chunk_off[k], m = p->chunk_off[k + 1] - o0;
                if (int e = stage_interp(*ch, dcp + o0, fw, 1)) return e;
                p->launches_exec += ch->launches_exec; ch->launches_exec = 0;
                CFB_CUDA_OK(cudaEventRecord(p->chunk_ev[k], st));
                CFB_CUDA_OK(cudaStreamWaitEvent(cs, p->chunk_ev[k], 0));
                CFB_CUDA_OK(cudaMemcpyAsync(c + o0, dcp + o0, (size_t)m * sizeof(C), cudaMemcpyDeviceToHost, cs));
            }
            CFB_CUDA_OK(cudaStreamSynchronize(cs));
        }
        CFB_CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    }
    if (p->type == 1) { if (nc) CFB_CUDA_OK(cudaMemcpyAsync(dc.p, c, nc, cudaMemcpyHostToDevice, st)); }
    else CFB_CUDA_OK(cudaMemcpyAsync(dfk.p, fk, nk, cudaMemcpyHostToDevice, st));
    if (int e = execute(p, dc.as<C>(), dfk.as<C>())) return e;
    if (p->type == 1) CFB_CUDA_OK(cudaMemcpyAsync(fk, dfk.p, nk, cudaMemcpyDeviceToHost, st));
    else if (nc) CFB_CUDA_OK(cudaMemcpyAsync(c, dc.p, nc, cudaMemcpyDeviceToHost, st));
    CFB_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

template <typename T>
static int get_ints(Plan<T> *p, int what, int *out)
{
    if (!p || !out) return CFB_ERR_BAD_ARG;
    DeviceGuard guard(p->device);
    CFB_CUDA_OK(cudaStreamSynchronize(p->stream));
    int total = 0;
    if (p->M >= 0 && p->sorted) CFB_CUDA_OK(cudaMemcpy(&total, p->scalars.p, sizeof(int), cudaMemcpyDeviceToHost));
    const size_t nb = (size_t)p->nbins;
    auto d2h = [&](const DevBuf &b, size_t n) {
        return n == 0 ? cudaSuccess : cudaMemcpy(out, b.p, n * sizeof(int), cudaMemcpyDeviceToHost);
    };
    switch (what) {
        case 0: {
            int g[16] = {p->dim, p->nf1, p->nf2, p->nf3, p->ns, p->nbin[0], p->nbin[1], p->nbin[2], p->bs[0], p->bs[1], p->bs[2],
                         p->maxbatch, p->M, total, p->method, p->nbins};
            memcpy(out, g, sizeof(g));
            return 0;
        }
        case 1: CFB_CUDA_OK(d2h(p->binsize, nb)); return 0;
        case 2: CFB_CUDA_OK(d2h(p->binstartpts, nb)); return 0;
        case 3: CFB_CUDA_OK(d2h(p->numsubprob, nb)); return 0;
        case 4: CFB_CUDA_OK(d2h(p->subprobstartpts, nb + 1)); return 0;
        case 5: CFB_CUDA_OK(d2h(p->subprob_to_bin, (size_t)total)); return 0;
        case 6:
            if (int e = materialize_idxnupts(*p)) return e;
            CFB_CUDA_OK(cudaStreamSynchronize(p->stream));
            CFB_CUDA_OK(d2h(p->idxnupts, (size_t)(p->M > 0 ? p->M : 0)));
            return 0;
        default: return CFB_ERR_BAD_ARG;
    }
}

template <typename T>
static int get_reals(Plan<T> *p, int d, T *out)
{
    if (!p || !out) return CFB_ERR_BAD_ARG;
    DeviceGuard guard(p->device);
    if (d == -1) { out[0] = p->es_beta; out[1] = p->es_c; out[2] = p->es_halfwidth; return 0; }
    if (d < 0 || d >= p->dim || !p->fwker[d].p) return CFB_ERR_BAD_ARG;
    const int nf[3] = {p->nf1, p->nf2, p->nf3_global()};
    CFB_CUDA_OK(cudaStreamSynchronize(p->stream));
    CFB_CUDA_OK(cudaMemcpy(out, p->fwker[d].p, (size_t)(nf[d] / 2 + 1) * sizeof(T), cudaMemcpyDeviceToHost));
    return 0;
}

template <typename T>
static int get_timing(Plan<T> *p, float *out)
{
    if (!p || !out || !p->timing || !p->ev[4]) return CFB_ERR_BAD_ARG;
    DeviceGuard guard(p->device);
    CFB_CUDA_OK(cudaEventSynchronize(p->ev[4]));
    float memset_ms, a, fft, b, tot;
    cudaEventElapsedTime(&memset_ms, p->ev[0], p->ev[1]);
    cudaEventElapsedTime(&a, p->ev[1], p->ev[2]);
    cudaEventElapsedTime(&fft, p->ev[2], p->ev[3]);
    cudaEventElapsedTime(&b, p->ev[3], p->ev[4]);
    cudaEventElapsedTime(&tot, p->ev[0], p->ev[4]);
    out[0] = p->type == 1 ? a : b;    // spread | interp
    out[1] = fft;
    out[2] = p->type == 1 ? b : a;    // deconvolve | amplify
    out[3] = memset_ms;
    out[4] = tot;
    return 0;
}

// plan-time numbers without a device (tests of the host logic run where no GPU exists)
template <typename T>
static int host_params(int type, int dim, const int *nmodes, double tol, const cufinufft_opts *opts, int *oi, double *od)
{
    Plan<T> p;
    int ier = plan_host_setup<T>(&p, type, dim, nmodes, 1, 1, (T)tol, 1, opts);
    if (ier) return ier;
    const int v[16] = {p.ns, p.nf1, p.nf2, p.nf3, p.bs[0], p.bs[1], p.bs[2], p.nbin[0], p.nbin[1], p.nbin[2], p.nbins,
                       p.method, p.sorted ? 1 : 0, p.maxbatch, p.opts.gpu_method, (int)(2 + 3.0 * (T)(p.ns / 2.0))};
    memcpy(oi, v, sizeof(v));
    od[0] = (double)p.es_beta; od[1] = (double)p.es_c; od[2] = (double)p.es_halfwidth;
    return 0;
}

// what setpts would choose for M points, without a device: internal bins, tile, work-item size
// (choose_internal_bins, spread.cu) -- for the host-logic tests
template <typename T>
static int host_workplan(int type, int dim, const int *nmodes, double tol, const cufinufft_opts *opts, long long M, int *o)
{
    Plan<T> p;
    int ier = plan_host_setup<T>(&p, type, dim, nmodes, 1, 1, (T)tol, 1, opts);
    if (ier) return ier;
    p.M = (int)(M > 2147483647LL ? 2147483647LL : M);
    plan_tile_geometry(p);
    choose_internal_bins(p, M);
    const int v[16] = {p.ibs[0], p.ibs[1], p.ibs[2], p.spb[0], p.spb[1], p.spb[2], p.nibins, p.nbins, p.imaxsub, p.ilist ? 1 : 0,
                       p.tile_cells, p.tile_sy, p.tile_sz, p.sm_warps, p.tile_pad, p.tile_cost};
    memcpy(o, v, sizeof(v));
    return 0;
}

template <typename T>
static int stage_only(Plan<T> *p, typename Plan<T>::C *c, typename Plan<T>::C *fw, int nt, bool spread)
{
    using C = typename Plan<T>::C;
    if (!p || p->M < 0 || nt < 1) return CFB_ERR_BAD_ARG;
    DeviceGuard guard(p->device);
    p->launches_exec = 0;
    if (spread) {
        CFB_CUDA_OK(cudaMemsetAsync(fw, 0, (size_t)nt * p->grid_cells() * sizeof(C), p->stream));
        return stage_spread(*p, c, fw, nt);
    }
    return stage_interp(*p, c, fw, nt);
}

// ---- z-slab plans (slab.cu) ----------------------------------------------------
template <typename T>
static int slab_info(Plan<T> *p, long long *o)
{
    if (!p || !o || !p->slab) return CFB_ERR_BAD_ARG;
    DeviceGuard guard(p->device);
    int outside = 0;
    if (p->M >= 0) {
        CFB_CUDA_OK(cudaStreamSynchronize(p->stream));
        CFB_CUDA_OK(cudaMemcpy(&outside, p->scalars.template as<int>() + 3, sizeof(int), cudaMemcpyDeviceToHost));
    }
    const long long v[12] = {p->z0, p->z1, p->tile_pad, p->nf3, p->nf1, p->nf2, p->nf3g, (long long)p->nf1 * p->nf2,
                             p->slab_rank, p->slab_world, outside, p->ns};
    memcpy(o, v, sizeof(v));
    return 0;
}

template <typename T, typename F>
static int slab_call(Plan<T> *p, int want_type, F &&fn)
{
    if (!p || !p->slab || (want_type && p->type != want_type)) return CFB_ERR_BAD_ARG;
    if (p->M < 0) return CFB_ERR_NO_POINTS_SET;
    DeviceGuard guard(p->device);
    return fn(*p);
}

}  // namespace cfb

// =============================== C ABI ==========================================
using cfb::Plan;

#define PD(h) ((h) ? (h)->p : nullptr)

extern "C" {

const char *cufinufft_b200_version(void) { return "cufinufft-b200 0.1 (API of cuFINUFFT 1.3)"; }

int cufinufft_b200_host_params(int type, int dim, const int *nmodes, double tol, int single_precision,
                               const cufinufft_opts *opts, int *out_ints16, double *out_reals3)
{
    if (!out_ints16 || !out_reals3) return CFB_ERR_BAD_ARG;
    return single_precision ? cfb::host_params<float>(type, dim, nmodes, tol, opts, out_ints16, out_reals3)
                            : cfb::host_params<double>(type, dim, nmodes, tol, opts, out_ints16, out_reals3);
}
int cufinufft_b200_host_workplan(int type, int dim, const int *nmodes, double tol, int single_precision,
                                 const cufinufft_opts *opts, long long M, int *out_ints16)
{
    if (!out_ints16) return CFB_ERR_BAD_ARG;
    return single_precision ? cfb::host_workplan<float>(type, dim, nmodes, tol, opts, M, out_ints16)
                            : cfb::host_workplan<double>(type, dim, nmodes, tol, opts, M, out_ints16);
}
int cufinufft_b200_phihat_quadrature(int nf, int ns, double beta, double es_c, double halfwidth, int single_precision,
                                     void *f, double *a_reim)
{
    if (!f || !a_reim || ns < 2 || ns > cfb::MAX_NS || nf < 2) return CFB_ERR_BAD_ARG;
    if (single_precision) cfb::fseries_precomp<float>(nf, ns, (float)beta, (float)es_c, (float)halfwidth, (float *)f, a_reim);
    else cfb::fseries_precomp<double>(nf, ns, beta, es_c, halfwidth, (double *)f, a_reim);
    return 0;
}

int cufinufft_default_opts(int type, int dim, cufinufft_opts *opts) { return cfb::default_opts(type, dim, opts); }
int cufinufftf_default_opts(int type, int dim, cufinufft_opts *opts) { return cfb::default_opts(type, dim, opts); }

int cufinufft_makeplan(int type, int dim, int *nmodes, int iflag, int ntransf, double tol, int maxbatchsize,
                       cufinufft_plan *plan, cufinufft_opts *opts)
{
    if (!plan) return CFB_ERR_BAD_ARG;
    *plan = nullptr;
    Plan<double> *p = nullptr;
    int ier = cfb::makeplan<double>(type, dim, nmodes, iflag, ntransf, tol, maxbatchsize, &p, opts);
    if (ier) return ier;
    *plan = new cufinufft_plan_s{p, {}, {}};
    return 0;
}
int cufinufftf_makeplan(int type, int dim, int *nmodes, int iflag, int ntransf, float tol, int maxbatchsize,
                        cufinufftf_plan *plan, cufinufft_opts *opts)
{
    if (!plan) return CFB_ERR_BAD_ARG;
    *plan = nullptr;
    Plan<float> *p = nullptr;
    int ier = cfb::makeplan<float>(type, dim, nmodes, iflag, ntransf, tol, maxbatchsize, &p, opts);
    if (ier) return ier;
    *plan = new cufinufftf_plan_s{p, {}, {}};
    return 0;
}

int cufinufft_setpts(int M, double *x, double *y, double *z, int, double *, double *, double *, cufinufft_plan plan)
{ return cfb::setpts<double>(PD(plan), M, x, y, z); }
int cufinufftf_setpts(int M, float *x, float *y, float *z, int, float *, float *, float *, cufinufftf_plan plan)
{ return cfb::setpts<float>(PD(plan), M, x, y, z); }

int cufinufft_execute(cuDoubleComplex *c, cuDoubleComplex *fk, cufinufft_plan plan)
{ return cfb::execute<double>(PD(plan), reinterpret_cast<double2 *>(c), reinterpret_cast<double2 *>(fk)); }
int cufinufftf_execute(cuFloatComplex *c, cuFloatComplex *fk, cufinufftf_plan plan)
{ return cfb::execute<float>(PD(plan), reinterpret_cast<float2 *>(c), reinterpret_cast<float2 *>(fk)); }

int cufinufft_destroy(cufinufft_plan plan)
{
    if (!plan) return 1;
    int dev = plan->p ? plan->p->device : 0;
    { cfb::DeviceGuard g(dev); if (plan->p) cudaStreamSynchronize(plan->p->stream); plan->dc.release(); plan->dfk.release(); }
    int ier = cfb::destroy(plan->p);
    delete plan;
    return ier;
}
int cufinufftf_destroy(cufinufftf_plan plan)
{
    if (!plan) return 1;
    int dev = plan->p ? plan->p->device : 0;
    { cfb::DeviceGuard g(dev); if (plan->p) cudaStreamSynchronize(plan->p->stream); plan->dc.release(); plan->dfk.release(); }
    int ier = cfb::destroy(plan->p);
    delete plan;
    return ier;
}

int cufinufft_set_stream(cufinufft_plan plan, void *s) { if (!PD(plan)) return CFB_ERR_BAD_ARG; plan->p->stream = (cudaStream_t)s; return 0; }
int cufinufftf_set_stream(cufinufftf_plan plan, void *s) { if (!PD(plan)) return CFB_ERR_BAD_ARG; plan->p->stream = (cudaStream_t)s; return 0; }

int cufinufft_setpts_host(int M, const double *x, const double *y, const double *z, cufinufft_plan plan)
{ return cfb::setpts_host<double>(PD(plan), M, x, y, z); }
int cufinufftf_setpts_host(int M, const float *x, const float *y, const float *z, cufinufftf_plan plan)
{ return cfb::setpts_host<float>(PD(plan), M, x, y, z); }
int cufinufft_execute_host(cuDoubleComplex *c, cuDoubleComplex *fk, cufinufft_plan plan)
{ if (!PD(plan)) return CFB_ERR_BAD_ARG; return cfb::execute_host<double>(plan->p, reinterpret_cast<double2 *>(c), reinterpret_cast<double2 *>(fk), plan->dc, plan->dfk); }
int cufinufftf_execute_host(cuFloatComplex *c, cuFloatComplex *fk, cufinufftf_plan plan)
{ if (!PD(plan)) return CFB_ERR_BAD_ARG; return cfb::execute_host<float>(plan->p, reinterpret_cast<float2 *>(c), reinterpret_cast<float2 *>(fk), plan->dc, plan->dfk); }

int cufinufft_spread(cuDoubleComplex *c, cuDoubleComplex *fw, int nt, cufinufft_plan plan)
{ return cfb::stage_only<double>(PD(plan), reinterpret_cast<double2 *>(c), reinterpret_cast<double2 *>(fw), nt, true); }
int cufinufftf_spread(cuFloatComplex *c, cuFloatComplex *fw, int nt, cufinufftf_plan plan)
{ return cfb::stage_only<float>(PD(plan), reinterpret_cast<float2 *>(c), reinterpret_cast<float2 *>(fw), nt, true); }
int cufinufft_interp(cuDoubleComplex *c, cuDoubleComplex *fw, int nt, cufinufft_plan plan)
{ return cfb::stage_only<double>(PD(plan), reinterpret_cast<double2 *>(c), reinterpret_cast<double2 *>(fw), nt, false); }
int cufinufftf_interp(cuFloatComplex *c, cuFloatComplex *fw, int nt, cufinufftf_plan plan)
{ return cfb::stage_only<float>(PD(plan), reinterpret_cast<float2 *>(c), reinterpret_cast<float2 *>(fw), nt, false); }

int cufinufft_get_ints(cufinufft_plan plan, int what, int *out) { return cfb::get_ints<double>(PD(plan), what, out); }
int cufinufftf_get_ints(cufinufftf_plan plan, int what, int *out) { return cfb::get_ints<float>(PD(plan), what, out); }
int cufinufft_get_reals(cufinufft_plan plan, int d, double *out) { return cfb::get_reals<double>(PD(plan), d, out); }
int cufinufftf_get_reals(cufinufftf_plan plan, int d, float *out) { return cfb::get_reals<float>(PD(plan), d, out); }

int cufinufft_set_timing(cufinufft_plan plan, int on) { if (!PD(plan)) return CFB_ERR_BAD_ARG; plan->p->timing = on != 0; return 0; }
int cufinufftf_set_timing(cufinufftf_plan plan, int on) { if (!PD(plan)) return CFB_ERR_BAD_ARG; plan->p->timing = on != 0; return 0; }
int cufinufft_get_timing(cufinufft_plan plan, float *out) { return cfb::get_timing<double>(PD(plan), out); }
int cufinufftf_get_timing(cufinufftf_plan plan, float *out) { return cfb::get_timing<float>(PD(plan), out); }
int cufinufft_set_interp_engine(cufinufft_plan plan, int e) { if (!PD(plan) || e < 0 || e > 2) return CFB_ERR_BAD_ARG; plan->p->interp_engine = e; return 0; }
int cufinufftf_set_interp_engine(cufinufftf_plan plan, int e) { if (!PD(plan) || e < 0 || e > 2) return CFB_ERR_BAD_ARG; plan->p->interp_engine = e; return 0; }
int cufinufft_set_sort_levels(cufinufft_plan plan, int l) { if (!PD(plan) || l < 0 || (l & 3) > 2 || ((l >> 2) & 3) > 2 || l > 31) return CFB_ERR_BAD_ARG; plan->p->sort_levels = l & 3; plan->p->sort_partition = (l >> 2) & 3; plan->p->key_generic = (l & 16) != 0; return 0; }
int cufinufftf_set_sort_levels(cufinufftf_plan plan, int l) { if (!PD(plan) || l < 0 || (l & 3) > 2 || ((l >> 2) & 3) > 2 || l > 31) return CFB_ERR_BAD_ARG; plan->p->sort_levels = l & 3; plan->p->sort_partition = (l >> 2) & 3; plan->p->key_generic = (l & 16) != 0; return 0; }
int cufinufft_get_launch_counts(cufinufft_plan plan, int *o) { if (!PD(plan) || !o) return CFB_ERR_BAD_ARG; o[0] = plan->p->launches_setpts; o[1] = plan->p->launches_exec; return 0; }
int cufinufftf_get_launch_counts(cufinufftf_plan plan, int *o) { if (!PD(plan) || !o) return CFB_ERR_BAD_ARG; o[0] = plan->p->launches_setpts; o[1] = plan->p->launches_exec; return 0; }

// ---- z-slab decomposition (include/cufinufft_b200.h) ----
int cufinufft_slab_makeplan(int type, int *nmodes, int iflag, double tol, int rank, int world, cufinufft_plan *plan, cufinufft_opts *opts)
{
    if (!plan) return CFB_ERR_BAD_ARG;
    *plan = nullptr;
    if (world < 1) return CFB_ERR_BAD_ARG;
    Plan<double> *p = nullptr;
    int ier = cfb::makeplan<double>(type, 3, nmodes, iflag, 1, tol, 1, &p, opts, rank, world);
    if (ier) return ier;
    *plan = new cufinufft_plan_s{p, {}, {}};
    return 0;
}
int cufinufftf_slab_makeplan(int type, int *nmodes, int iflag, float tol, int rank, int world, cufinufftf_plan *plan, cufinufft_opts *opts)
{
    if (!plan) return CFB_ERR_BAD_ARG;
    *plan = nullptr;
    if (world < 1) return CFB_ERR_BAD_ARG;
    Plan<float> *p = nullptr;
    int ier = cfb::makeplan<float>(type, 3, nmodes, iflag, 1, tol, 1, &p, opts, rank, world);
    if (ier) return ier;
    *plan = new cufinufftf_plan_s{p, {}, {}};
    return 0;
}
int cufinufft_slab_info(cufinufft_plan plan, long long *out12) { return cfb::slab_info<double>(PD(plan), out12); }
int cufinufftf_slab_info(cufinufftf_plan plan, long long *out12) { return cfb::slab_info<float>(PD(plan), out12); }

int cufinufft_slab_type2(cuDoubleComplex *c, cuDoubleComplex *fk, cufinufft_plan plan)
{ return cfb::slab_call<double>(PD(plan), 2, [&](Plan<double> &p) { return cfb::slab_type2<double>(p, reinterpret_cast<double2 *>(c), reinterpret_cast<const double2 *>(fk)); }); }
int cufinufftf_slab_type2(cuFloatComplex *c, cuFloatComplex *fk, cufinufftf_plan plan)
{ return cfb::slab_call<float>(PD(plan), 2, [&](Plan<float> &p) { return cfb::slab_type2<float>(p, reinterpret_cast<float2 *>(c), reinterpret_cast<const float2 *>(fk)); }); }

int cufinufft_slab_type1_spread(cuDoubleComplex *c, cufinufft_plan plan)
{ return cfb::slab_call<double>(PD(plan), 1, [&](Plan<double> &p) { return cfb::slab_type1_spread<double>(p, reinterpret_cast<const double2 *>(c)); }); }
int cufinufftf_slab_type1_spread(cuFloatComplex *c, cufinufftf_plan plan)
{ return cfb::slab_call<float>(PD(plan), 1, [&](Plan<float> &p) { return cfb::slab_type1_spread<float>(p, reinterpret_cast<const float2 *>(c)); }); }

int cufinufft_slab_halo_pack(int side, cuDoubleComplex *buf, cufinufft_plan plan)
{ if (side < 0 || side > 1 || !buf) return CFB_ERR_BAD_ARG;
  return cfb::slab_call<double>(PD(plan), 0, [&](Plan<double> &p) { return cfb::slab_halo_pack<double>(p, side, reinterpret_cast<double2 *>(buf)); }); }
int cufinufftf_slab_halo_pack(int side, cuFloatComplex *buf, cufinufftf_plan plan)
{ if (side < 0 || side > 1 || !buf) return CFB_ERR_BAD_ARG;
  return cfb::slab_call<float>(PD(plan), 0, [&](Plan<float> &p) { return cfb::slab_halo_pack<float>(p, side, reinterpret_cast<float2 *>(buf)); }); }
int cufinufft_slab_halo_add(int side, cuDoubleComplex *buf, cufinufft_plan plan)
{ if (side < 0 || side > 1 || !buf) return CFB_ERR_BAD_ARG;
  return cfb::slab_call<double>(PD(plan), 0, [&](Plan<double> &p) { return cfb::slab_halo_add<double>(p, side, reinterpret_cast<const double2 *>(buf)); }); }
int cufinufftf_slab_halo_add(int side, cuFloatComplex *buf, cufinufftf_plan plan)
{ if (side < 0 || side > 1 || !buf) return CFB_ERR_BAD_ARG;
  return cfb::slab_call<float>(PD(plan), 0, [&](Plan<float> &p) { return cfb::slab_halo_add<float>(p, side, reinterpret_cast<const float2 *>(buf)); }); }

int cufinufft_slab_type1_finish(cuDoubleComplex *fk_partial, cufinufft_plan plan)
{ return cfb::slab_call<double>(PD(plan), 1, [&](Plan<double> &p) { return cfb::slab_type1_finish<double>(p, reinterpret_cast<double2 *>(fk_partial)); }); }
int cufinufftf_slab_type1_finish(cuFloatComplex *fk_partial, cufinufftf_plan plan)
{ return cfb::slab_call<float>(PD(plan), 1, [&](Plan<float> &p) { return cfb::slab_type1_finish<float>(p, reinterpret_cast<float2 *>(fk_partial)); }); }

}  // extern "C"
