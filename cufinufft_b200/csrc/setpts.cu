// setpts.cu -- the setpts hot path: bin histogram + in-bin rank, fused scans and
// subproblem map, inverse permutation + bin-ordered rescaled coordinates.
//
// What the reference does (src/2d/spread2d_wrapper.cu:386-613, kernels
// CalcBinSize_noghost_* / CalcInvertofGlobalSortIdx_* in src/{1,2,3}d/spreadinterp*.cu,
// CalcSubProb_* / MapBintoSubProb_* in src/precision_independent.cu:35-127, two
// Thrust scans, one blocking D2H of totalnumsubprob and a cudaMalloc inside setpts):
// the same RESULTS are produced here --
//   binsize, binstartpts, numsubprob, subprobstartpts, subprob_to_bin: bit-exact,
//   idxnupts: a bin-major permutation (within-bin order is a race in the reference
//   too, src/2d/spreadinterp2d.cu:120-121) --
// with a different schedule, all stream-ordered with no host sync:
//   K1 bin_count      warp-aggregated (match_any) histogram atomics -> rank per point
//   K2 scan_bins      ONE kernel: exclusive scan of counts, integer ceil-div subproblem
//                     counts, their inclusive scan and totalnumsubprob (device scalar)
//   K3 map_subprob    one thread per subproblem slot, binary search in subprobstartpts
//   K4 place_points   idxnupts + physically permuted, already-rescaled coordinates
//                     (xs,ys,zs) so spread/interp read them coalesced and never
//                     re-evaluate RESCALE (the reference recomputes it 3x per point).
// Algorithmic bytes per point: K1 d*sF + 4, K4 d*sF + 4 + 4 + d*sF  (DESIGN.md).
#include "cfb_device.cuh"

namespace cfb {

template <typename T, int DIM>
__device__ __forceinline__ int point_bin(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                                         int i, int nf1, int nf2, int nf3, int bs1, int bs2, int bs3,
                                         int nb1, int nb2, int nb3, T &xr, T &yr, T &zr)
{
    xr = rescale(x[i], nf1);
    int b = bin_coord(xr, bs1, nb1);
    if (DIM > 1) { yr = rescale(y[i], nf2); b += nb1 * bin_coord(yr, bs2, nb2); }
    if (DIM > 2) { zr = rescale(z[i], nf3); b += nb1 * nb2 * bin_coord(zr, bs3, nb3); }
    return b;
}

// K1: histogram + rank.  Lanes of a warp that fall in the same bin are aggregated into
// one atomicAdd (clustered inputs put most of a warp in one bin).
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
bin_count_kernel(int M, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                 int nf1, int nf2, int nf3, int bs1, int bs2, int bs3, int nb1, int nb2, int nb3,
                 int *__restrict__ binsize, int *__restrict__ sortidx)
{
    const int lane = threadIdx.x & 31;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < M; base += (long long)gridDim.x * blockDim.x) {
        int i = (int)(base + threadIdx.x);
        bool valid = i < M;
        T xr, yr, zr;
        int b = valid ? point_bin<T, DIM>(x, y, z, i, nf1, nf2, nf3, bs1, bs2, bs3, nb1, nb2, nb3, xr, yr, zr) : -1 - lane;
        unsigned peers = __match_any_sync(0xffffffffu, b);
        int leader = __ffs(peers) - 1;
        int rank_in_group = __popc(peers & ((1u << lane) - 1));
        int basecnt = 0;
        if (valid && lane == leader) basecnt = atomicAdd(&binsize[b], __popc(peers));
        basecnt = __shfl_sync(0xffffffffu, basecnt, leader);
        if (valid) sortidx[i] = basecnt + rank_in_group;
    }
}

// K2: one block walks the bins in chunks of blockDim.x with a running carry.  Produces
// binstartpts (exclusive), numsubprob = ceil(binsize/maxsub) (integer: the reference's float
// ceil drops the last subproblem above 2^24 points per bin, precision_independent.cu:71),
// subprobstartpts (inclusive scan with leading 0) and scalars[0] = totalnumsubprob.
__global__ void __launch_bounds__(1024)
scan_bins_kernel(int nbins, int maxsub, const int *__restrict__ binsize, int *__restrict__ binstartpts,
                 int *__restrict__ numsubprob, int *__restrict__ subprobstartpts, int *__restrict__ scalars)
{
    __shared__ int wsum_a[32], wsum_b[32];
    __shared__ int carry_a, carry_b;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) { carry_a = 0; carry_b = 0; subprobstartpts[0] = 0; }
    __syncthreads();
    for (int base = 0; base < nbins; base += blockDim.x) {
        int b = base + threadIdx.x;
        int cnt = b < nbins ? binsize[b] : 0;
        int nsp = (cnt + maxsub - 1) / maxsub;
        int sa = cnt, sb = nsp;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int ta = __shfl_up_sync(0xffffffffu, sa, o), tb = __shfl_up_sync(0xffffffffu, sb, o);
            if (lane >= o) { sa += ta; sb += tb; }
        }
        if (lane == 31) { wsum_a[wid] = sa; wsum_b[wid] = sb; }
        __syncthreads();
        if (wid == 0) {
            int va = lane < nw ? wsum_a[lane] : 0, vb = lane < nw ? wsum_b[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int ta = __shfl_up_sync(0xffffffffu, va, o), tb = __shfl_up_sync(0xffffffffu, vb, o);
                if (lane >= o) { va += ta; vb += tb; }
            }
            wsum_a[lane] = va; wsum_b[lane] = vb;   // inclusive over warps
        }
        __syncthreads();
        int offa = carry_a + (wid ? wsum_a[wid - 1] : 0);
        int offb = carry_b + (wid ? wsum_b[wid - 1] : 0);
        if (b < nbins) {
            binstartpts[b] = offa + sa - cnt;
            numsubprob[b] = nsp;
            subprobstartpts[b + 1] = offb + sb;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) { carry_a = offa + sa; carry_b = offb + sb; }
        __syncthreads();
    }
    if (threadIdx.x == 0) scalars[0] = carry_b;
}

// K3: subprob_to_bin[s] = the bin whose slot range contains s (upper bound launch: slots
// beyond totalnumsubprob exit).  Replaces MapBintoSubProb_* + the blocking D2H/cudaMalloc.
__global__ void __launch_bounds__(256)
map_subprob_kernel(int nbins, int maxslots, const int *__restrict__ subprobstartpts,
                   const int *__restrict__ scalars, int *__restrict__ subprob_to_bin)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= maxslots || s >= scalars[0]) return;
    int lo = 0, hi = nbins;                 // find largest b with subprobstartpts[b] <= s
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (subprobstartpts[mid] <= s) lo = mid; else hi = mid;
    }
    subprob_to_bin[s] = lo;
}

// K4: inverse permutation and bin-ordered rescaled coordinates.
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
place_points_kernel(int M, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                    int nf1, int nf2, int nf3, int bs1, int bs2, int bs3, int nb1, int nb2, int nb3,
                    const int *__restrict__ binstartpts, const int *__restrict__ sortidx,
                    int *__restrict__ idxnupts, T *__restrict__ xs, T *__restrict__ ys, T *__restrict__ zs)
{
    for (long long ii = (long long)blockIdx.x * blockDim.x + threadIdx.x; ii < M; ii += (long long)gridDim.x * blockDim.x) {
        int i = (int)ii;
        T xr, yr, zr;
        int b = point_bin<T, DIM>(x, y, z, i, nf1, nf2, nf3, bs1, bs2, bs3, nb1, nb2, nb3, xr, yr, zr);
        int pos = binstartpts[b] + sortidx[i];
        idxnupts[pos] = i;
        xs[pos] = xr;
        if (DIM > 1) ys[pos] = yr;
        if (DIM > 2) zs[pos] = zr;
    }
}

// gpu_sort = 0 (GM): identity order (TrivialGlobalSortIdx_*, src/precision_independent.cu:57),
// still with rescaled coordinates stored once.
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
trivial_order_kernel(int M, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                     int nf1, int nf2, int nf3, int *__restrict__ idxnupts, T *__restrict__ xs,
                     T *__restrict__ ys, T *__restrict__ zs)
{
    for (long long ii = (long long)blockIdx.x * blockDim.x + threadIdx.x; ii < M; ii += (long long)gridDim.x * blockDim.x) {
        int i = (int)ii;
        idxnupts[i] = i;
        xs[i] = rescale(x[i], nf1);
        if (DIM > 1) ys[i] = rescale(y[i], nf2);
        if (DIM > 2) zs[i] = rescale(z[i], nf3);
    }
}

template <typename T, int DIM>
static int setpts_dim(Plan<T> &p)
{
    const int M = p.M;
    cudaStream_t st = p.stream;
    const T *x = p.kx, *y = p.ky, *z = p.kz;
    int *binsize = p.binsize.template as<int>();
    int *binstart = p.binstartpts.template as<int>();
    int *nsub = p.numsubprob.template as<int>();
    int *substart = p.subprobstartpts.template as<int>();
    int *s2b = p.subprob_to_bin.template as<int>();
    int *scal = p.scalars.template as<int>();
    int *sortidx = p.sortidx.template as<int>();
    int *idx = p.idxnupts.template as<int>();
    T *xs = p.xs.template as<T>(), *ys = p.ys.template as<T>(), *zs = p.zs.template as<T>();
    const int threads = 256;
    long long want = ((long long)M + threads - 1) / threads;
    int blocks = (int)(want < 1 ? 1 : (want > (long long)p.num_sms * 32 ? (long long)p.num_sms * 32 : want));
    p.launches_setpts = 0;

    if (!p.sorted) {
        if (M > 0) {
            trivial_order_kernel<T, DIM><<<blocks, threads, 0, st>>>(M, x, y, z, p.nf1, p.nf2, p.nf3, idx, xs, ys, zs);
            p.launches_setpts++;
        }
        CFB_CUDA_OK(cudaMemsetAsync(scal, 0, 8 * sizeof(int), st));
        CFB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    CFB_CUDA_OK(cudaMemsetAsync(binsize, 0, sizeof(int) * (size_t)p.nbins, st));
    if (M > 0) {
        bin_count_kernel<T, DIM><<<blocks, threads, 0, st>>>(M, x, y, z, p.nf1, p.nf2, p.nf3, p.bs[0], p.bs[1], p.bs[2],
                                                           p.nbin[0], p.nbin[1], p.nbin[2], binsize, sortidx);
        p.launches_setpts++;
    }
    scan_bins_kernel<<<1, 1024, 0, st>>>(p.nbins, p.opts.gpu_maxsubprobsize, binsize, binstart, nsub, substart, scal);
    p.launches_setpts++;
    int maxslots = p.nbins + M / p.opts.gpu_maxsubprobsize + 1;
    map_subprob_kernel<<<(maxslots + 255) / 256, 256, 0, st>>>(p.nbins, maxslots, substart, scal, s2b);
    p.launches_setpts++;
    if (M > 0) {
        place_points_kernel<T, DIM><<<blocks, threads, 0, st>>>(M, x, y, z, p.nf1, p.nf2, p.nf3, p.bs[0], p.bs[1], p.bs[2],
                                                              p.nbin[0], p.nbin[1], p.nbin[2], binstart, sortidx, idx, xs, ys, zs);
        p.launches_setpts++;
    }
    CFB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T>
int stage_setpts(Plan<T> &p)
{
    const size_t M = (size_t)(p.M > 0 ? p.M : 1);
    CFB_CUDA_OK(p.sortidx.reserve(M * sizeof(int)));
    CFB_CUDA_OK(p.idxnupts.reserve(M * sizeof(int)));
    CFB_CUDA_OK(p.xs.reserve(M * sizeof(T)));
    if (p.dim > 1) CFB_CUDA_OK(p.ys.reserve(M * sizeof(T)));
    if (p.dim > 2) CFB_CUDA_OK(p.zs.reserve(M * sizeof(T)));
    size_t maxslots = (size_t)p.nbins + M / (size_t)p.opts.gpu_maxsubprobsize + 1;
    CFB_CUDA_OK(p.subprob_to_bin.reserve(maxslots * sizeof(int)));
    switch (p.dim) {
        case 1: return setpts_dim<T, 1>(p);
        case 2: return setpts_dim<T, 2>(p);
        default: return setpts_dim<T, 3>(p);
    }
}
template int stage_setpts<float>(Plan<float> &);
template int stage_setpts<double>(Plan<double> &);

}  // namespace cfb
