// setpts.cu -- the setpts hot path: one counting sort of the points by (bin, stencil cell),
// bin counts / offsets and the subproblem map derived from it, and the bin-ordered point
// records the spread / interp kernels stream.
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
//   K1 key_count     key = bin * cells_per_bin + stencil cell inside the bin; warp-aggregated
//                    (match_any) histogram atomics return the rank of each point in its key
//   K2 scan_reduce / scan_top / scan_apply   three-phase exclusive scan of the key counts
//   K3 ref_bins     binsize[b] = off[(b+1)*cpb] - off[b*cpb], binstartpts[b] = off[b*cpb], integer
//                    ceil-div subproblem counts
//   K4 the same three-phase scan over the subproblem counts -> subprobstartpts, totalnumsubprob
//                    (device scalar)
//   K5 map_subprob   one thread per subproblem slot, binary search in subprobstartpts
//   K6 place_points  ONE 16-byte (fp32) / 32-byte (fp64) record per point
//                    {x_rescaled, y_rescaled, z_rescaled, original index} scattered to its sorted
//                    slot: a single sector write per point instead of four, and the spread /
//                    interp kernels read it back with one coalesced vector load and never
//                    re-evaluate RESCALE (the reference recomputes it 3x per point per execute).
// Sorting INSIDE the bin by stencil cell (the reference leaves that order to a race) makes
// consecutive points share their whole stencil; the spread / interp kernels exploit that by
// keeping a run of such points in registers (spreadinterp.cuh).  The fine histogram is used
// when it is not much larger than the point set (nkeys <= 8 M + 2^22), otherwise the key is
// the bin alone.
// Algorithmic bytes per point: K1 d*sF + 4, K6 d*sF + 4 + 4 + rec  (DESIGN.md).
#include "cfb_device.cuh"

namespace cfb {

// reference bin b, sub-bin s inside it and stencil-origin cell inside the sub-bin for one coordinate
template <typename T>
__device__ __forceinline__ void dim_key(T xr, int d, const SortGeo &g, int &b, int &s, int &cell)
{
    b = bin_coord(xr, g.bs[d], g.nb[d]);
    int origin = b * g.bs[d];
    s = 0;
    if (g.spb[d] > 1) {
        s = (int)floor((xr - (T)origin) / (T)g.ibs[d]);
        s = s < 0 ? 0 : (s >= g.spb[d] ? g.spb[d] - 1 : s);
        origin += s * g.ibs[d];
    }
    if (g.bankc > 0) {
        // bank-class order: offset of the first stencil cell inside the bin's tile (origin - halo), clamped exactly
        // as the tile interpolation kernel clamps it (spreadinterp.cuh: interp_tile_kernel)
        const int e = d == 0 ? g.bex : (d == 1 ? g.bey : g.bez);
        const int o = stencil_start(xr, g.ns) - (origin - g.bpad);
        cell = o < 0 ? 0 : (o > e - g.ns ? e - g.ns : o);
        return;
    }
    cell = g.nk[d] > 1 ? stencil_cell(xr, g.ns, origin, g.nk[d]) : 0;
}

// sort key = ((reference bin * sub-bins per bin + sub-bin) * cells per sub-bin + stencil cell)
template <typename T, int DIM>
__device__ __forceinline__ int point_key(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                                         int i, const SortGeo &g, T &xr, T &yr, T &zr, int *count_outside = nullptr)
{
    int b, sb, c;
    const int cs1 = g.bankc > 0 ? g.bex : g.nk[0], cs2 = g.bankc > 0 ? g.bex * g.bey : g.nk[0] * g.nk[1];   // cell strides
    xr = rescale(x[i], g.nf[0]);
    dim_key(xr, 0, g, b, sb, c);
    int bin = b, sub = sb, cell = c;
    if (DIM > 1) {
        yr = rescale(y[i], g.nf[1]);
        dim_key(yr, 1, g, b, sb, c);
        bin += g.nb[0] * b; sub += g.spb[0] * sb; cell += cs1 * c;
    }
    if (DIM > 2) {
        zr = rescale(z[i], g.nfz);
        // slab plans: bins over the slab-local planes; the record keeps the GLOBAL z_r (weights are
        // then bit-identical to the undivided transform), the kernels shift the stencil start.
        // A point outside the slab (caller error) is pulled onto its edge and counted.  The stencil
        // of a point stays inside the halo for zl in [zlo - 1/2, zhi], so a caller whose own
        // rescale differs from ours in the last bit at a slab boundary is still served exactly.
        T zl = zr - (T)g.zshift;
        if (zl < (T)g.zlo - (T)0.5 || zl > (T)g.zhi) {
            if (count_outside) atomicAdd(count_outside, 1);
            zr = zl < (T)g.zlo ? (T)(g.zlo + g.zshift) : (T)(g.zhi + g.zshift);
            zl = zr - (T)g.zshift;
        }
        dim_key(zl, 2, g, b, sb, c);
        bin += g.nb[0] * g.nb[1] * b; sub += g.spb[0] * g.spb[1] * sb; cell += cs2 * c;
    }
    if (g.bankc > 0) cell &= g.bankc - 1;            // the bank class (bankc is a power of two)
    return (bin * g.spbt + sub) * g.cpb + cell;
}

// K1: histogram + rank.  Lanes of a warp that fall on the same key are aggregated into
// one atomicAdd (clustered inputs put most of a warp on one key).
// Four points per thread and iteration: the twelve coordinate loads are in flight together, then the
// four histogram atomics, then the four rank stores (one point per iteration left the kernel at the
// latency of load -> atomic -> store chains: 1.7 TB/s with every table L2-resident, profiles/r01zi).
constexpr int SP_UNROLL = 4;

template <typename T, int DIM>
__global__ void __launch_bounds__(256)
key_count_kernel(int M, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                 const SortGeo g, int *__restrict__ keycnt, int *__restrict__ rank, int *__restrict__ outside)
{
    const int lane = threadIdx.x & 31;
    const long long span = (long long)blockDim.x * SP_UNROLL;
    for (long long base = (long long)blockIdx.x * span; base < M; base += (long long)gridDim.x * span) {
        int k[SP_UNROLL];
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u) {
            const long long i = base + (long long)u * blockDim.x + threadIdx.x;
            T xr, yr, zr;
            k[u] = i < M ? point_key<T, DIM>(x, y, z, (int)i, g, xr, yr, zr, outside) : -1 - lane;
        }
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u) {
            const long long i = base + (long long)u * blockDim.x + threadIdx.x;
            const bool valid = i < M;
            const unsigned peers = __match_any_sync(0xffffffffu, k[u]);
            const int leader = __ffs(peers) - 1;
            const int rank_in_group = __popc(peers & ((1u << lane) - 1));
            int basecnt = 0;
            if (valid && lane == leader) basecnt = atomicAdd(&keycnt[k[u]], __popc(peers));
            basecnt = __shfl_sync(0xffffffffu, basecnt, leader);
            if (valid) rank[i] = basecnt + rank_in_group;
        }
    }
}

// ---- three-phase exclusive scan of n ints, in place, tiles of SCAN_TILE per block ----------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_PER_THREAD = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ int block_exclusive_scan(int v, int *wsum /*[32]*/, int &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    if (lane == 31) wsum[wid] = s;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        wsum[lane] = w;
    }
    __syncthreads();
    total = wsum[nw - 1];
    int excl = s - v + (wid ? wsum[wid - 1] : 0);
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(long long n, const int *__restrict__ a, int *__restrict__ tilesum)
{
    __shared__ int wsum[32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_PER_THREAD;
    int s = 0;
    if (base + SCAN_PER_THREAD <= n) {
        const int4 *p = reinterpret_cast<const int4 *>(a + base);
#pragma unroll
        for (int j = 0; j < SCAN_PER_THREAD / 4; ++j) { int4 v = p[j]; s += v.x + v.y + v.z + v.w; }
    } else {
        for (int j = 0; j < SCAN_PER_THREAD; ++j) if (base + j < n) s += a[base + j];
    }
    int total;
    block_exclusive_scan(s, wsum, total);
    if (threadIdx.x == 0) tilesum[blockIdx.x] = total;
}

// single block: exclusive scan of the tile sums (ntiles <= ~2^18 for 2^31 keys)
__global__ void __launch_bounds__(1024)
scan_top_kernel(int ntiles, int *__restrict__ tilesum)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < ntiles ? tilesum[i] : 0;
        int total;
        const int excl = block_exclusive_scan(v, wsum, total);
        const int c = carry;
        if (i < ntiles) tilesum[i] = c + excl;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(long long n, int *__restrict__ a, const int *__restrict__ tilesum)
{
    __shared__ int wsum[32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_PER_THREAD;
    int v[SCAN_PER_THREAD];
    int s = 0;
    const bool full = base + SCAN_PER_THREAD <= n;
    if (full) {
        const int4 *p = reinterpret_cast<const int4 *>(a + base);
#pragma unroll
        for (int j = 0; j < SCAN_PER_THREAD / 4; ++j) {
            int4 q = p[j];
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_PER_THREAD; ++j) v[j] = base + j < n ? a[base + j] : 0;
    }
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j) s += v[j];
    int total;
    int run = block_exclusive_scan(s, wsum, total) + tilesum[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j) { const int t = v[j]; v[j] = run; run += t; }
    if (full) {
        int4 *p = reinterpret_cast<int4 *>(a + base);
#pragma unroll
        for (int j = 0; j < SCAN_PER_THREAD / 4; ++j) p[j] = make_int4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_PER_THREAD; ++j) if (base + j < n) a[base + j] = v[j];
    }
}

// K3+K4 (multi-block): per reference bin its size, its start (= the key offset), its subproblem count
// (integer ceil: the reference's float ceil drops the last subproblem above 2^24 points per bin,
// precision_independent.cu:71); the counts are then scanned in place into subprobstartpts.
__global__ void __launch_bounds__(256)
ref_bins_kernel(int nbins, int cpb, int maxsub, const int *__restrict__ keyoff, int *__restrict__ binsize,
                int *__restrict__ binstartpts, int *__restrict__ numsubprob, int *__restrict__ subprobstartpts)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbins) {
        const int p0 = keyoff[(size_t)b * cpb], cnt = keyoff[(size_t)(b + 1) * cpb] - p0;
        const int nsp = (cnt + maxsub - 1) / maxsub;
        binsize[b] = cnt; binstartpts[b] = p0; numsubprob[b] = nsp; subprobstartpts[b] = nsp;
    } else if (b == nbins) subprobstartpts[b] = 0;
}

// internal bins (finer than the reference's): subproblem count per internal bin, to be scanned
// in place (entry nibins = 0 becomes the total)
__global__ void __launch_bounds__(256)
isub_count_kernel(int nibins, int cpb, int maxsub, const int *__restrict__ keyoff, int *__restrict__ isubstart)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nibins) isubstart[b] = (keyoff[(size_t)(b + 1) * cpb] - keyoff[(size_t)b * cpb] + maxsub - 1) / maxsub;
    else if (b == nibins) isubstart[b] = 0;
}

// K5: subprob_to_bin[s] = the bin whose slot range contains s (upper bound launch: slots
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

template <typename T>
__device__ __forceinline__ void store_rec(PtRec<T> *dst, T xr, T yr, T zr, int i);
template <>
__device__ __forceinline__ void store_rec<float>(PtRec<float> *dst, float xr, float yr, float zr, int i)
{
    *reinterpret_cast<float4 *>(dst) = make_float4(xr, yr, zr, __int_as_float(i));
}
template <>
__device__ __forceinline__ void store_rec<double>(PtRec<double> *dst, double xr, double yr, double zr, int i)
{
    double2 *d = reinterpret_cast<double2 *>(dst);
    d[0] = make_double2(xr, yr);
    d[1] = make_double2(zr, __longlong_as_double((long long)i));
}

// K6: scatter one record per point to its sorted slot (four points per thread and iteration:
// coordinate and rank loads first, then the offset gathers, then the record stores).
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
place_points_kernel(int M, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                    const SortGeo g, const int *__restrict__ keyoff, const int *__restrict__ rank,
                    PtRec<T> *__restrict__ recs)
{
    const long long span = (long long)blockDim.x * SP_UNROLL;
    for (long long base = (long long)blockIdx.x * span; base < M; base += (long long)gridDim.x * span) {
        T xr[SP_UNROLL], yr[SP_UNROLL], zr[SP_UNROLL];
        int k[SP_UNROLL], rk[SP_UNROLL];
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u) {
            const long long i = base + (long long)u * blockDim.x + threadIdx.x;
            xr[u] = 0; yr[u] = 0; zr[u] = 0; k[u] = 0; rk[u] = 0;
            if (i < M) {
                k[u] = point_key<T, DIM>(x, y, z, (int)i, g, xr[u], yr[u], zr[u]);
                rk[u] = rank[i];
            }
        }
        int pos[SP_UNROLL];
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u) pos[u] = keyoff[k[u]] + rk[u];
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u) {
            const long long i = base + (long long)u * blockDim.x + threadIdx.x;
            if (i < M) store_rec<T>(recs + pos[u], xr[u], yr[u], zr[u], (int)i);
        }
    }
}

// gpu_sort = 0 (GM): identity order (TrivialGlobalSortIdx_*, src/precision_independent.cu:57),
// still with rescaled coordinates stored once.
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
trivial_order_kernel(int M, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                     const SortGeo g, PtRec<T> *__restrict__ recs, int *__restrict__ outside)
{
    for (long long ii = (long long)blockIdx.x * blockDim.x + threadIdx.x; ii < M; ii += (long long)gridDim.x * blockDim.x) {
        int i = (int)ii;
        T xr, yr = 0, zr = 0;
        point_key<T, DIM>(x, y, z, i, g, xr, yr, zr, outside);      // rescale (+ slab clamp); key unused
        store_rec<T>(recs + i, xr, yr, zr, i);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
extract_idx_kernel(int M, const PtRec<T> *__restrict__ recs, int *__restrict__ idx)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x)
        idx[i] = rec_index(recs[i]);
}

// ---- second level of the two-level order -----------------------------------------------------
// When the (bin, stencil cell) histogram gets much larger than L2 (config 5 on one GPU: 4.3 GB)
// its random atomics and gathers run at DRAM-sector speed.  The global pass then sorts by
// internal bin only (a table of a few MB: L2-resident) and this kernel orders every work item
// (<= maxsub consecutive points of one bin) by stencil cell: records in registers (8 per thread),
// cell histogram + ranks with shared-memory atomics, block scan, write back inside the item's own
// range.  All global traffic of this pass is a coalesced read and a write confined to a 64 KB window.
struct LocalSortArgs {
    const int *keyoff, *s2b, *substart, *nsub;
    int maxsub;
    int nb1, nb2, rbs[3], ibs[3], spb1, spb2, spbt;
    int ns, nkf[3], cpbf, zshift;
};

constexpr int LS_THREADS = 256;
constexpr int LS_RPT = 8;                       // records per thread
constexpr int LS_MAXKEYS = 4096;

template <typename T, int DIM>
__global__ void __launch_bounds__(LS_THREADS)
local_sort_kernel(const LocalSortArgs a, PtRec<T> *__restrict__ recs)
{
    __shared__ int cnt[LS_MAXKEYS];
    __shared__ int wsum[32];
    const int nsub = *a.nsub;
    for (int s = blockIdx.x; s < nsub; s += gridDim.x) {
        const int bin = a.s2b[s];
        const int k = s - a.substart[bin];
        const int p0 = a.keyoff[bin], p1 = a.keyoff[bin + 1];
        const int pstart = p0 + k * a.maxsub;
        const int n = min(a.maxsub, p1 - pstart);
        if (n < 2) continue;
        const int rb = bin / a.spbt, sub = bin - rb * a.spbt;
        const int b1 = rb % a.nb1, b23 = rb / a.nb1, s1 = sub % a.spb1, s23 = sub / a.spb1;
        const int o1 = b1 * a.rbs[0] + s1 * a.ibs[0];
        const int o2 = DIM > 1 ? (b23 % a.nb2) * a.rbs[1] + (s23 % a.spb2) * a.ibs[1] : 0;
        const int o3 = DIM > 2 ? (b23 / a.nb2) * a.rbs[2] + (s23 / a.spb2) * a.ibs[2] : 0;
        for (int c0 = 0; c0 < n; c0 += LS_THREADS * LS_RPT) {
            const int m = min(LS_THREADS * LS_RPT, n - c0);
            PtRec<T> *base = recs + pstart + c0;
            for (int i = threadIdx.x; i < a.cpbf; i += LS_THREADS) cnt[i] = 0;
            __syncthreads();
            PtRec<T> r[LS_RPT];
            int kr[LS_RPT];                                   // key << 13 | rank inside the key
#pragma unroll
            for (int j = 0; j < LS_RPT; ++j) {
                const int i = j * LS_THREADS + threadIdx.x;
                kr[j] = -1;
                if (i < m) {
                    r[j] = base[i];
                    int key = a.nkf[0] > 1 ? stencil_cell(r[j].x, a.ns, o1, a.nkf[0]) : 0;
                    if (DIM > 1 && a.nkf[1] > 1) key += a.nkf[0] * stencil_cell(r[j].y, a.ns, o2, a.nkf[1]);
                    if (DIM > 2 && a.nkf[2] > 1) key += a.nkf[0] * a.nkf[1] * stencil_cell(r[j].z - (T)a.zshift, a.ns, o3, a.nkf[2]);
                    kr[j] = (key << 13) | atomicAdd(&cnt[key], 1);       // (warp-aggregating this atomic was measured: slower)
                }
            }
            __syncthreads();
            // exclusive scan of cnt[0 .. cpbf): 16 entries per thread + block scan of the partial sums
            {
                constexpr int PER = LS_MAXKEYS / LS_THREADS;
                int v[PER], sum = 0;
#pragma unroll
                for (int j = 0; j < PER; ++j) { const int i = threadIdx.x * PER + j; v[j] = i < a.cpbf ? cnt[i] : 0; sum += v[j]; }
                int total;
                int run = block_exclusive_scan(sum, wsum, total);
#pragma unroll
                for (int j = 0; j < PER; ++j) { const int i = threadIdx.x * PER + j; if (i < a.cpbf) cnt[i] = run; run += v[j]; }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < LS_RPT; ++j)
                if (kr[j] >= 0) base[cnt[kr[j] >> 13] + (kr[j] & 8191)] = r[j];
            __syncthreads();
        }
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
    int *rank = p.sortidx.template as<int>();
    int *keyoff = p.keyoff.template as<int>();
    int *tilesum = p.tilesum.template as<int>();
    PtRec<T> *recs = p.recs.template as<PtRec<T>>();
    const int threads = 256;
    long long want = ((long long)M + threads - 1) / threads;
    int blocks = (int)(want < 1 ? 1 : (want > (long long)p.num_sms * 32 ? (long long)p.num_sms * 32 : want));
    p.launches_setpts = 0;

    CFB_CUDA_OK(cudaMemsetAsync(scal, 0, 8 * sizeof(int), st));
    if (!p.sorted) {
        if (M > 0) {
            trivial_order_kernel<T, DIM><<<blocks, threads, 0, st>>>(M, x, y, z, p.sortgeo, recs, p.slab ? scal + 3 : nullptr);
            p.launches_setpts++;
        }
        CFB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    const SortGeo g = p.sortgeo;
    const long long nkeys = (long long)p.nibins * g.cpb;
    const long long nscan = nkeys + 1;                       // trailing total
    const int ntiles = (int)((nscan + SCAN_TILE - 1) / SCAN_TILE);
    CFB_CUDA_OK(cudaMemsetAsync(keyoff, 0, sizeof(int) * (size_t)nscan, st));
    if (M > 0) {
        key_count_kernel<T, DIM><<<blocks, threads, 0, st>>>(M, x, y, z, g, keyoff, rank, p.slab ? scal + 3 : nullptr);
        p.launches_setpts++;
    }
    scan_reduce_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(nscan, keyoff, tilesum);
    scan_top_kernel<<<1, 1024, 0, st>>>(ntiles, tilesum);
    scan_apply_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(nscan, keyoff, tilesum);
    {   // reference-facing arrays from the key offsets: binsize, binstartpts (= the offsets themselves),
        // numsubprob (integer ceil), subprobstartpts by the same three-phase scan, total -> scalars[0]
        const long long nrs = (long long)p.nbins + 1;
        const int rtiles = (int)((nrs + SCAN_TILE - 1) / SCAN_TILE);
        ref_bins_kernel<<<(int)((nrs + 255) / 256), 256, 0, st>>>(p.nbins, g.cpb * g.spbt, p.opts.gpu_maxsubprobsize, keyoff, binsize,
                                                                 binstart, nsub, substart);
        scan_reduce_kernel<<<rtiles, SCAN_THREADS, 0, st>>>(nrs, substart, tilesum);
        scan_top_kernel<<<1, 1024, 0, st>>>(rtiles, tilesum);
        scan_apply_kernel<<<rtiles, SCAN_THREADS, 0, st>>>(nrs, substart, tilesum);
        CFB_CUDA_OK(cudaMemcpyAsync(scal, substart + p.nbins, sizeof(int), cudaMemcpyDeviceToDevice, st));
    }
    p.launches_setpts += 7;
    int maxslots = p.nbins + M / p.opts.gpu_maxsubprobsize + 1;
    map_subprob_kernel<<<(maxslots + 255) / 256, 256, 0, st>>>(p.nbins, maxslots, substart, scal, s2b);
    p.launches_setpts++;
    if (p.ilist) {
        // the tile engines' own work list over the internal bins (same construction, finer bins / larger items)
        int *isub = p.isubstart.template as<int>();
        const long long nis = (long long)p.nibins + 1;
        const int itiles = (int)((nis + SCAN_TILE - 1) / SCAN_TILE);
        isub_count_kernel<<<(int)((nis + 255) / 256), 256, 0, st>>>(p.nibins, g.cpb, p.imaxsub, keyoff, isub);
        scan_reduce_kernel<<<itiles, SCAN_THREADS, 0, st>>>(nis, isub, tilesum);
        scan_top_kernel<<<1, 1024, 0, st>>>(itiles, tilesum);
        scan_apply_kernel<<<itiles, SCAN_THREADS, 0, st>>>(nis, isub, tilesum);
        int islots = p.nibins + M / p.imaxsub + 1;
        map_subprob_kernel<<<(islots + 255) / 256, 256, 0, st>>>(p.nibins, islots, isub, isub + p.nibins, p.is2b.template as<int>());
        p.launches_setpts += 5;
    }
    if (M > 0) {
        place_points_kernel<T, DIM><<<blocks, threads, 0, st>>>(M, x, y, z, g, keyoff, rank, recs);
        p.launches_setpts++;
    }
    if (M > 0 && p.local_sort) {
        LocalSortArgs la;
        la.keyoff = keyoff;
        la.s2b = p.ilist ? p.is2b.template as<int>() : s2b;
        la.substart = p.ilist ? p.isubstart.template as<int>() : substart;
        la.nsub = p.ilist ? p.isubstart.template as<int>() + p.nibins : scal;
        la.maxsub = p.ilist ? p.imaxsub : p.opts.gpu_maxsubprobsize;
        la.nb1 = p.nbin[0]; la.nb2 = p.nbin[1];
        for (int d = 0; d < 3; ++d) { la.rbs[d] = p.bs[d]; la.ibs[d] = p.ibs[d]; la.nkf[d] = g.nkf[d]; }
        la.spb1 = p.spb[0]; la.spb2 = p.spb[1]; la.spbt = g.spbt;
        la.ns = p.ns; la.cpbf = g.cpbf; la.zshift = g.zshift;
        local_sort_kernel<T, DIM><<<p.num_sms * 8, LS_THREADS, 0, st>>>(la, recs);
        p.launches_setpts++;
    }
    CFB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T>
int stage_setpts(Plan<T> &p)
{
    const size_t M = (size_t)(p.M > 0 ? p.M : 1);
    // key space: (bin, stencil cell in bin) when that histogram is not much larger than the
    // point set, else the bin alone
    SortGeo &g = p.sortgeo;
    const int nf[3] = {p.nf1, p.nf2, p.nf3};
    choose_internal_bins(p, (long long)p.M);        // p.ibs / p.spb / p.nibins + the tile geometry that goes with them
    g.ns = p.ns;
    g.nfz = p.nf3_global(); g.zshift = p.slab ? p.zshift : 0;
    g.zlo = p.slab ? p.tile_pad : -(1 << 30); g.zhi = p.slab ? p.tile_pad + (p.z1 - p.z0) : (1 << 30);   // ordinary plans: never clamp
    long long cpb = 1;
    for (int d = 0; d < 3; ++d) {
        g.nf[d] = nf[d]; g.bs[d] = p.bs[d]; g.nb[d] = p.nbin[d];
        g.ibs[d] = p.ibs[d]; g.spb[d] = p.spb[d];
        g.nk[d] = g.nkf[d] = d < p.dim ? p.ibs[d] + (p.ns & 1) : 1;
        cpb *= g.nk[d];
    }
    g.spbt = p.spb[0] * p.spb[1] * p.spb[2];
    g.cpbf = (int)(cpb < (1LL << 30) ? cpb : (1LL << 30));
    // type-2 plans whose points go to the tile interpolation engine: order the points of a bin by shared-memory
    // bank class instead of stencil cell (that engine is thread-per-point and has no use for runs; what it
    // needs is that the lanes sharing a wavefront hit different banks).  128 / sizeof(C) classes per bin.
    g.bankc = 0; g.bex = g.bey = g.bez = 1; g.bpad = p.tile_pad;
    // 3-D only: the 2-D kernels have few conflicts to begin with (neighbouring lanes sit in the same or the next
    // cell of a row) and measured 15-35 % slower in this order (configs 2 and 4, profiles/r02g).
    if (p.type == 2 && p.dim == 3 && p.sorted && p.fine_sort_allowed && interp_tile_applies(p)) {
        g.bankc = 128 / (int)sizeof(typename Plan<T>::C);
        g.bex = p.ibs[0] + 2 * p.tile_pad;
        g.bey = p.dim > 1 ? p.ibs[1] + 2 * p.tile_pad : 1;
        g.bez = p.dim > 2 ? p.ibs[2] + 2 * p.tile_pad : 1;
        cpb = g.bankc;
    }
    p.bank_classes = g.bankc;
    long long nkeys = cpb * p.nibins;
    // one level (global (bin, cell) histogram) while that table is comfortably L2-sized and not much
    // larger than the point set; two levels (bins globally, cells per work item) when it is not;
    // bins only when a bin has more stencil cells than the local sort handles
    const bool dense = nkeys <= 8LL * (long long)M + (1LL << 22);        // else runs have one point anyway: bins only
    bool one_level = p.fine_sort_allowed && dense && nkeys * 4 <= (2048LL << 20);
    bool two_level = p.fine_sort_allowed && dense && !one_level;
    if (p.sort_levels == 2) { one_level = false; two_level = p.fine_sort_allowed; }
    if (p.sort_levels == 1 && p.fine_sort_allowed && nkeys <= 2000000000LL) { one_level = true; two_level = false; }
    if (g.bankc > 0) { one_level = nkeys <= 2000000000LL; two_level = false; if (!one_level) { g.bankc = 0; p.bank_classes = 0; } }
    p.local_sort = two_level && p.sorted && cpb <= LS_MAXKEYS;
    if (!one_level) {
        g.nk[0] = g.nk[1] = g.nk[2] = 1;
        cpb = 1;
        nkeys = p.nibins;
    }
    g.cpb = (int)cpb;
    CFB_CUDA_OK(p.recs.reserve(M * sizeof(PtRec<T>)));
    if (p.sorted) {
        CFB_CUDA_OK(p.sortidx.reserve(M * sizeof(int)));
        CFB_CUDA_OK(p.keyoff.reserve(((size_t)nkeys + 1 + 4) * sizeof(int)));
        CFB_CUDA_OK(p.tilesum.reserve(((size_t)(nkeys + 1) / SCAN_TILE + 2) * sizeof(int)));
    }
    size_t maxslots = (size_t)p.nbins + M / (size_t)p.opts.gpu_maxsubprobsize + 1;
    CFB_CUDA_OK(p.subprob_to_bin.reserve(maxslots * sizeof(int)));
    if (p.ilist) {
        CFB_CUDA_OK(p.isubstart.reserve(((size_t)p.nibins + 1 + 4) * sizeof(int)));
        CFB_CUDA_OK(p.is2b.reserve(((size_t)p.nibins + M / (size_t)p.imaxsub + 1) * sizeof(int)));
    }
    p.idx_valid = false;
    switch (p.dim) {
        case 1: return setpts_dim<T, 1>(p);
        case 2: return setpts_dim<T, 2>(p);
        default: return setpts_dim<T, 3>(p);
    }
}

// idxnupts is materialised only when somebody asks for it (plan introspection)
template <typename T>
int materialize_idxnupts(Plan<T> &p)
{
    if (p.idx_valid || p.M <= 0) return 0;
    CFB_CUDA_OK(p.idxnupts.reserve((size_t)p.M * sizeof(int)));
    extract_idx_kernel<T><<<p.num_sms * 8, 256, 0, p.stream>>>(p.M, p.recs.template as<PtRec<T>>(), p.idxnupts.template as<int>());
    CFB_CUDA_OK(cudaGetLastError());
    p.idx_valid = true;
    return 0;
}

template int stage_setpts<float>(Plan<float> &);
template int stage_setpts<double>(Plan<double> &);
template int materialize_idxnupts<float>(Plan<float> &);
template int materialize_idxnupts<double>(Plan<double> &);

}  // namespace cfb
