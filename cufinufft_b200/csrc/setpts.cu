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
//   K0 coarse_count / coarse_scatter   (only when the key table is far beyond L2) the points are first dealt
//                    into a few hundred buckets = ranges of bin rows, so that K1 and K6 work in L2-sized windows
//   K1 key_count     key = bin * cells_per_bin + stencil cell inside the bin; one reduction (no return
//                    value) per distinct key of a warp into the key table
//   K2 scan_reduce / scan_top / scan_apply   three-phase exclusive scan of the key counts
//   K3 ref_bins     binsize[b] = off[(b+1)*cpb] - off[b*cpb], binstartpts[b] = off[b*cpb], integer
//                    ceil-div subproblem counts
//   K4 the same three-phase scan over the subproblem counts -> subprobstartpts, totalnumsubprob
//                    (device scalar)
//   K5 map_subprob   one thread per subproblem slot, binary search in subprobstartpts
//   K6 place_points  ONE 16-byte (fp32) / 32-byte (fp64) record per point
//                    {x_rescaled, y_rescaled, z_rescaled, original index} scattered to its sorted
//                    slot, which comes from the key's cursor (the scanned table entry, advanced by one
//                    warp-aggregated atomic): a single sector write per point instead of four, and the spread /
//                    interp kernels read it back with one coalesced vector load and never
//                    re-evaluate RESCALE (the reference recomputes it 3x per point per execute).
// Sorting INSIDE the bin by stencil cell (the reference leaves that order to a race) makes
// consecutive points share their whole stencil; the spread / interp kernels exploit that by
// keeping a run of such points in registers (spreadinterp.cuh).  The fine histogram is used
// when it is not much larger than the point set (nkeys <= 8 M + 2^22), otherwise the key is
// the bin alone.
// Algorithmic bytes per point: K1 d*sF, K6 d*sF + rec; with K0: + d*sF + (d*sF + rec), K1 and K6 read rec instead
// of d*sF.
#include <cstdlib>
#include "cfb_device.cuh"

namespace cfb {

// reference bin b, sub-bin s inside it and stencil-origin cell inside the sub-bin for one coordinate
template <typename T>
__device__ __forceinline__ void dim_key(T xr, int d, const SortGeo &g, int &b, int &s, int &cell)
{
    int origin;
    if (g.lg_ibs[d] >= 0) {
        // power-of-two sizes: q = internal-bin coordinate before the clamps; floor(x_r / bs) = q >> lg(spb) is the
        // reference's own value (division by a power of two is exact), and so is the sub-bin (x_r - origin is exact)
        const int q = (int)floor(xr * (T)g.inv_ibs[d]);
        b = q >> g.lg_spb[d];
        b = b >= g.nb[d] ? b - 1 : b;
        b = b < 0 ? 0 : b;
        s = q - (b << g.lg_spb[d]);
        s = s < 0 ? 0 : (s >= g.spb[d] ? g.spb[d] - 1 : s);
        origin = ((b << g.lg_spb[d]) + s) << g.lg_ibs[d];
    } else {
        b = bin_coord(xr, g.bs[d], g.nb[d]);
        origin = b * g.bs[d];
        s = 0;
        if (g.spb[d] > 1) {
            s = (int)floor((xr - (T)origin) / (T)g.ibs[d]);
            s = s < 0 ? 0 : (s >= g.spb[d] ? g.spb[d] - 1 : s);
            origin += s * g.ibs[d];
        }
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

// RESCALE of one point (+ the slab clamp of z): what the point record keeps
template <typename T, int DIM, bool WITH_X = true>
__device__ __forceinline__ void rescale_point(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z, long long i,
                                              const SortGeo &g, T &xr, T &yr, T &zr, int *count_outside = nullptr)
{
    xr = WITH_X ? rescale(x[i], g.nf[0]) : (T)0;
    yr = 0; zr = 0;
    if (DIM > 1) yr = rescale(y[i], g.nf[1]);
    if (DIM > 2) {
        zr = rescale(z[i], g.nfz);
        // slab plans: bins over the slab-local planes; the record keeps the GLOBAL z_r (weights are
        // then bit-identical to the undivided transform), the kernels shift the stencil start.
        // A point outside the slab (caller error) is pulled onto its edge and counted.  The stencil
        // of a point stays inside the halo for zl in [zlo - 1/2, zhi], so a caller whose own
        // rescale differs from ours in the last bit at a slab boundary is still served exactly.
        const T zl = zr - (T)g.zshift;
        if (zl < (T)g.zlo - (T)0.5 || zl > (T)g.zhi) {
            if (count_outside) atomicAdd(count_outside, 1);
            zr = zl < (T)g.zlo ? (T)(g.zlo + g.zshift) : (T)(g.zhi + g.zshift);
        }
    }
}

// sort key = ((reference bin * sub-bins per bin + sub-bin) * cells per sub-bin + stencil cell) of a rescaled point
template <typename T, int DIM>
__device__ __forceinline__ int key_of(T xr, T yr, T zr, const SortGeo &g)
{
    int b, sb, c;
    const int cs1 = g.bankc > 0 ? g.bex : g.nk[0], cs2 = g.bankc > 0 ? g.bex * g.bey : g.nk[0] * g.nk[1];   // cell strides
    dim_key(xr, 0, g, b, sb, c);
    int bin = b, sub = sb, cell = c;
    if (DIM > 1) {
        dim_key(yr, 1, g, b, sb, c);
        bin += g.nb[0] * b; sub += g.spb[0] * sb; cell += cs1 * c;
    }
    if (DIM > 2) {
        dim_key(zr - (T)g.zshift, 2, g, b, sb, c);
        bin += g.nb[0] * g.nb[1] * b; sub += g.spb[0] * g.spb[1] * sb; cell += cs2 * c;
    }
    if (g.bankc > 0) cell &= g.bankc - 1;            // the bank class (bankc is a power of two)
    return (bin * g.spbt + sub) * g.cpb + cell;
}

// The same key, specialised at compile time for power-of-two bin sizes (the defaults) and for what the key holds
// inside a bin: KM = 1 the bin alone, 2 the stencil cell, 3 the bank class.  The generic version above reads ~40
// kernel parameters behind run-time branches -- 195 instructions per point, the constant loads alone kept the
// address-divergence unit 90 % busy (profiles/r02n) -- this one needs ~90 and no branch.  KM = 0: generic.
template <typename T, int KM>
__device__ __forceinline__ void dim_key_p2(T xr, int d, const SortGeo &g, int &iq, int &cell)
{
    const int q = (int)floor(xr * (T)g.inv_ibs[d]);
    int b = q >> g.lg_spb[d];
    b = b >= g.nb[d] ? b - 1 : b;
    b = b < 0 ? 0 : b;
    int s = q - (b << g.lg_spb[d]);
    s = s < 0 ? 0 : (s >= g.spb[d] ? g.spb[d] - 1 : s);
    iq = (b << g.lg_spb[d]) + s;                       // internal-bin coordinate; b = iq >> lg_spb, s = iq & (spb - 1)
    cell = 0;
    if (KM == 2) {
        int k = stencil_start(xr, g.ns) - ((iq << g.lg_ibs[d]) - g.ns / 2) - ((g.ns & 1) ? 0 : 1);
        cell = k < 0 ? 0 : (k >= g.nk[d] ? g.nk[d] - 1 : k);
    } else if (KM == 3) {
        const int e = d == 0 ? g.bex : (d == 1 ? g.bey : g.bez);
        const int o = stencil_start(xr, g.ns) - ((iq << g.lg_ibs[d]) - g.bpad);
        cell = o < 0 ? 0 : (o > e - g.ns ? e - g.ns : o);
    }
}

template <typename T, int DIM, int KM>
__device__ __forceinline__ int key_any(T xr, T yr, T zr, const SortGeo &g)
{
    if (KM == 0) return key_of<T, DIM>(xr, yr, zr, g);
    int iq, c;
    dim_key_p2<T, KM>(xr, 0, g, iq, c);
    int bin = iq >> g.lg_spb[0], sub = iq & (g.spb[0] - 1), cell = c;
    if (DIM > 1) {
        dim_key_p2<T, KM>(yr, 1, g, iq, c);
        bin += g.nb[0] * (iq >> g.lg_spb[1]); sub += g.spb[0] * (iq & (g.spb[1] - 1));
        cell += (KM == 3 ? g.bex : g.nk[0]) * c;
    }
    if (DIM > 2) {
        dim_key_p2<T, KM>(zr - (T)g.zshift, 2, g, iq, c);
        bin += g.nb[0] * g.nb[1] * (iq >> g.lg_spb[2]); sub += g.spb[0] * g.spb[1] * (iq & (g.spb[2] - 1));
        cell += (KM == 3 ? g.bex * g.bey : g.nk[0] * g.nk[1]) * c;
    }
    if (KM == 1) return bin * g.spbt + sub;
    if (KM == 3) cell &= g.bankc - 1;
    return (bin * g.spbt + sub) * g.cpb + cell;
}

// One point of a setpts pass: from the caller's arrays (RESCALE here) or from the coarse-partitioned
// records (already rescaled; `idx` = the caller's index).
template <typename T, int DIM, bool FROM_RECS, int KM>
__device__ __forceinline__ int fetch_point(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
                                           const PtRec<T> *__restrict__ src, long long i, const SortGeo &g,
                                           T &xr, T &yr, T &zr, int &idx, int *count_outside = nullptr)
{
    if (FROM_RECS) {
        const PtRec<T> r = load_rec(src + i);
        xr = r.x; yr = r.y; zr = r.z; idx = rec_index(r);
    } else {
        rescale_point<T, DIM>(x, y, z, i, g, xr, yr, zr, count_outside);
        idx = (int)i;
    }
    return key_any<T, DIM, KM>(xr, yr, zr, g);
}

// Which lanes of the warp hold the same key?  MATCH.ANY answers in one instruction, but it runs on the address
// divergence unit at ~2.5 cycles per lane -- as slow as the atomic it is meant to save (profiles/r02n: that unit
// 80-90 % busy in every kernel that matched or reduced once per point).  Votes are cheap:
//   peers_by_bits   exact peer mask of a key known to fit in `nbits` bits: one vote per bit
//   peers_by_hash   lanes are grouped by the key's low five bits; each group's first lane is its leader and
//                   the mask holds the lanes of the group whose key EQUALS the leader's.  A lane outside its
//                   group's mask (another key with the same low bits) gets only itself.  Every lane is in
//                   exactly one returned mask, equal keys of one group share it: enough to aggregate atomics.
__device__ __forceinline__ unsigned peers_by_bits(int key, int nbits)
{
    unsigned m = 0xffffffffu;
    for (int b = 0; b < nbits; ++b) {
        const unsigned v = __ballot_sync(0xffffffffu, (key >> b) & 1);
        m &= ((key >> b) & 1) ? v : ~v;
    }
    return m;
}
__device__ __forceinline__ unsigned peers_by_hash(int key, int lane)
{
    unsigned m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const unsigned v = __ballot_sync(0xffffffffu, (key >> b) & 1);
        m &= ((key >> b) & 1) ? v : ~v;
    }
    const int lead = __ffs(m) - 1;
    const int kl = __shfl_sync(0xffffffffu, key, lead);
    const unsigned same = __ballot_sync(0xffffffffu, key == kl) & m;       // lanes of my group that equal its leader
    return key == kl ? same : (1u << lane);
}

// K1: histogram.  Lanes of a warp that fall on the same key are aggregated into one reduction
// (clustered inputs put most of a warp on one key; same-address atomics serialise in L2).
// Four points per thread and iteration: the twelve coordinate loads are in flight together, then the
// four histogram reductions (no return value: nothing waits for L2).
constexpr int SP_UNROLL = 4;

template <typename T, int DIM, bool FROM_RECS, int KM>
__global__ void __launch_bounds__(256)
key_count_kernel(int M, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z, const PtRec<T> *__restrict__ src,
                 const SortGeo g, int *__restrict__ keycnt, int *__restrict__ outside)
{
    const int lane = threadIdx.x & 31;
    const long long span = (long long)blockDim.x * SP_UNROLL;
    for (long long base = (long long)blockIdx.x * span; base < M; base += (long long)gridDim.x * span) {
        int k[SP_UNROLL];
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u) {
            const long long i = base + (long long)u * blockDim.x + threadIdx.x;
            T xr, yr, zr;
            int idx;
            k[u] = i < M ? fetch_point<T, DIM, FROM_RECS, KM>(x, y, z, src, i, g, xr, yr, zr, idx, outside) : -1 - lane;
        }
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u) {
            const unsigned peers = peers_by_hash(k[u], lane);
            if (k[u] >= 0 && lane == __ffs(peers) - 1) atomicAdd(&keycnt[k[u]], __popc(peers));
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

// K6: scatter one record per point to its sorted slot.  The slot comes from the key's CURSOR: the scanned
// table entry, advanced by one warp-aggregated atomic per distinct key of the warp (no rank array to write in
// K1 and read back here; the order inside a key is the arrival order either way).  After this kernel entry k
// of the table holds the END of key k = the start of key k + 1: readers use the table shifted by one entry
// (Plan::key_offsets).  Four points per thread and iteration: loads, then the four atomics, then the stores.
template <typename T, int DIM, bool FROM_RECS, int KM>
__global__ void __launch_bounds__(256)
place_points_kernel(int M, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z, const PtRec<T> *__restrict__ src,
                    const SortGeo g, int *__restrict__ cursor, PtRec<T> *__restrict__ recs)
{
    const int lane = threadIdx.x & 31;
    const long long span = (long long)blockDim.x * SP_UNROLL;
    for (long long base = (long long)blockIdx.x * span; base < M; base += (long long)gridDim.x * span) {
        T xr[SP_UNROLL], yr[SP_UNROLL], zr[SP_UNROLL];
        int k[SP_UNROLL], id[SP_UNROLL], pos[SP_UNROLL];
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u) {
            const long long i = base + (long long)u * blockDim.x + threadIdx.x;
            xr[u] = 0; yr[u] = 0; zr[u] = 0; k[u] = -1 - lane; id[u] = 0;
            if (i < M) k[u] = fetch_point<T, DIM, FROM_RECS, KM>(x, y, z, src, i, g, xr[u], yr[u], zr[u], id[u]);
        }
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u) {
            const unsigned peers = peers_by_hash(k[u], lane);
            const int leader = __ffs(peers) - 1;
            int b0 = 0;
            if (k[u] >= 0 && lane == leader) b0 = atomicAdd(&cursor[k[u]], __popc(peers));
            pos[u] = __shfl_sync(0xffffffffu, b0, leader) + __popc(peers & ((1u << lane) - 1));
        }
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u)
            if (k[u] >= 0) store_rec<T>(recs + pos[u], xr[u], yr[u], zr[u], id[u]);
    }
}

// ---- coarse partition: the locality pass in front of the counting sort ---------------------------
// The counting sort above scatters every point once -- histogram atomic, offset gather, record store -- at an
// address that is random over the whole key table and the whole record array when the caller's points come in
// random order.  Beyond L2 size that is one DRAM page miss and one read-modify-write of a half-written sector
// per point (config 3: 5.2 ms for 1e8 16-byte records, 0.6 TB/s; profiles/r01y).  The remedy is to limit the
// number of open write frontiers: the points are first dealt into <= CP_MAXB coarse buckets, contiguous ranges
// of the SAME key (bucket = key >> kshift, a few MB of records each), and the counting sort then reads them in
// that order, so that its atomics, gathers and record stores stay inside an L2-resident window which L2 merges
// into full lines before they reach DRAM.
//   coarse_count    histogram of the buckets in warp-private shared-memory counters (no atomics: the lanes of a warp
//                   that share a bucket are found with match_any, their leader does a plain read-modify-write),
//                   summed over the warps and added to the global counts once per block
//   scan_top        exclusive scan of the <= 4097 bucket counts (cursors)
//   coarse_scatter  tiles of CP_THREADS x CP_PPT points in registers, ranked the same way; the per-warp counts are
//                   scanned per bucket, ONE cursor atomic per tile and bucket reserves the tile's run, and the
//                   records {x_r, y_r, z_r, idx} go to cursor + warp offset + rank -- a few hundred open frontiers
// What was measured on the way (config 3, profiles/r02j-k): ranking with one shared-memory atomic per point costs
// 2 cycles per lane and pass (1.1 + 3.6 ms); global atomics on the few hundred bucket counters serialise at ~10
// cycles each whatever the warp aggregation (16 ms per pass).
// The bin arrays do not change (same histogram); only the within-key order does, which is a race in the
// reference as well (setpts.cu header).
constexpr int CP_MAXB = 1024;
constexpr int CP_THREADS = 256;
constexpr int CP_WARPS = CP_THREADS / 32;
#ifndef CFB_CP_PPT
#define CFB_CP_PPT 8
#endif
#ifndef CFB_CP_BPS
#define CFB_CP_BPS 8
#endif
constexpr int CP_PPT = CFB_CP_PPT;     // points per thread and tile
constexpr int CP_BPS = CFB_CP_BPS;     // blocks per SM of the two coarse kernels

// rank of this lane's point among the points of its warp that went to bucket b so far (warp-private counters)
__device__ __forceinline__ int warp_private_rank(int *wc, int b, int lane, int nbits)
{
    const unsigned peers = peers_by_bits(b, nbits);
    const int leader = __ffs(peers) - 1;
    int c = 0;
    if (b >= 0 && lane == leader) { c = wc[b]; wc[b] = c + __popc(peers); }
    __syncwarp();
    return __shfl_sync(0xffffffffu, c, leader) + __popc(peers & ((1u << lane) - 1));
}

// Bucket of a point: contiguous ranges of the sort key that are cheap to find.  The key is bin-major with x
// fastest, so a ROW of reference bins along x (all sub-bins and stencil cells included) is a contiguous key range:
// bucket = (by + nby * bz) >> kshift needs neither x nor the sub-bins nor the stencil cell -- a third of the key's
// instructions, and the counting pass does not even load x.  1-D: bucket = reference bin >> kshift.
template <typename T, int KM>
__device__ __forceinline__ int ref_bin(T xr, int d, const SortGeo &g)
{
    if (KM == 0) return bin_coord(xr, g.bs[d], g.nb[d]);
    int b = (int)floor(xr * (T)g.inv_ibs[d]) >> g.lg_spb[d];
    b = b >= g.nb[d] ? b - 1 : b;
    return b < 0 ? 0 : b;
}
template <typename T, int DIM, int KM, bool WITH_X>
__device__ __forceinline__ int coarse_bucket(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z, long long i,
                                             const SortGeo &g, int kshift, T &xr, T &yr, T &zr, int *outside)
{
    rescale_point<T, DIM, WITH_X || DIM == 1>(x, y, z, i, g, xr, yr, zr, outside);
    if (DIM == 1) return ref_bin<T, KM>(xr, 0, g) >> kshift;
    int row = ref_bin<T, KM>(yr, 1, g);
    if (DIM > 2) row += g.nb[1] * ref_bin<T, KM>(zr - (T)g.zshift, 2, g);
    return row >> kshift;
}

// Both kernels give block `blk` the same contiguous range of tiles [blk * tpb, (blk + 1) * tpb): the count
// kernel writes the block's bucket counts into the matrix cnt[bucket][block]; its exclusive scan (bucket-major) is,
// for every block, where its points of every bucket start -- the scatter needs no global atomic at all (one cursor
// atomic per tile and bucket was tried first: atomics WITH a return value on a few hundred addresses serialise at
// ~100 ns each, 2.1 ms for config 3, profiles/r02s), and the order of the records is deterministic.
template <typename T, int DIM, int KM>
__global__ void __launch_bounds__(CP_THREADS)
coarse_count_kernel(int M, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z, const SortGeo g, int kshift,
                    int nb, int nbits, long long tpb, int *__restrict__ cnt, int *__restrict__ outside)
{
    extern __shared__ int cp_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int *wc = cp_smem + wid * nb;
    for (int i = threadIdx.x; i < CP_WARPS * nb; i += CP_THREADS) cp_smem[i] = 0;
    __syncthreads();
    const long long tile = (long long)CP_THREADS * CP_PPT;
    const long long p0 = (long long)blockIdx.x * tpb * tile;
    const long long p1 = p0 + tpb * tile < M ? p0 + tpb * tile : M;
    const long long span = (long long)CP_THREADS * SP_UNROLL;
    for (long long base = p0; base < p1; base += span) {
        int k[SP_UNROLL];
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u) {
            const long long i = base + (long long)u * CP_THREADS + threadIdx.x;
            T xr, yr, zr;
            k[u] = i < p1 ? coarse_bucket<T, DIM, KM, false>(x, y, z, i, g, kshift, xr, yr, zr, outside) : -1;
        }
#pragma unroll
        for (int u = 0; u < SP_UNROLL; ++u) {
            const unsigned peers = peers_by_bits(k[u], nbits);
            if (k[u] >= 0 && lane == __ffs(peers) - 1) wc[k[u]] += __popc(peers);
            __syncwarp();
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += CP_THREADS) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < CP_WARPS; ++w) t += cp_smem[w * nb + b];
        cnt[(size_t)b * gridDim.x + blockIdx.x] = t;
    }
}

template <typename T, int DIM, int KM>
__global__ void __launch_bounds__(CP_THREADS)
coarse_scatter_kernel(int M, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z, const SortGeo g, int kshift,
                      int nb, int nbits, long long tpb, const int *__restrict__ start, PtRec<T> *__restrict__ out)
{
    extern __shared__ int cp_smem[];                  // [CP_WARPS][nb] counts -> offsets, [nb] the tile's run starts, [nb] the block's cursors
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int *wc = cp_smem + wid * nb, *runstart = cp_smem + CP_WARPS * nb, *cursor = runstart + nb;
    for (int q = threadIdx.x; q < nb; q += CP_THREADS) cursor[q] = start[(size_t)q * gridDim.x + blockIdx.x];
    const long long tile = (long long)CP_THREADS * CP_PPT;
    const long long p0 = (long long)blockIdx.x * tpb * tile;
    const long long p1 = p0 + tpb * tile < M ? p0 + tpb * tile : M;
    for (long long base = p0; base < p1; base += tile) {
        for (int i = threadIdx.x; i < CP_WARPS * nb; i += CP_THREADS) cp_smem[i] = 0;
        __syncthreads();
        T xr[CP_PPT], yr[CP_PPT], zr[CP_PPT];
        int b[CP_PPT], r[CP_PPT];
#pragma unroll
        for (int u = 0; u < CP_PPT; ++u) {
            // a warp takes 32 x CP_PPT CONSECUTIVE points: coalesced loads, and a rank that follows the input order
            const long long i = base + (long long)wid * (32 * CP_PPT) + u * 32 + lane;
            xr[u] = 0; yr[u] = 0; zr[u] = 0; b[u] = -1;
            if (i < p1) b[u] = coarse_bucket<T, DIM, KM, true>(x, y, z, i, g, kshift, xr[u], yr[u], zr[u], nullptr);
        }
#pragma unroll
        for (int u = 0; u < CP_PPT; ++u) r[u] = warp_private_rank(wc, b[u], lane, nbits);
        __syncthreads();
        for (int q = threadIdx.x; q < nb; q += CP_THREADS) {       // thread q owns bucket q's cursor in every tile
            int t = 0;
#pragma unroll
            for (int w = 0; w < CP_WARPS; ++w) { const int c = cp_smem[w * nb + q]; cp_smem[w * nb + q] = t; t += c; }
            const int c0 = cursor[q];
            runstart[q] = c0; cursor[q] = c0 + t;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < CP_PPT; ++u) {
            const long long i = base + (long long)wid * (32 * CP_PPT) + u * 32 + lane;
            if (b[u] >= 0) store_rec<T>(out + (runstart[b[u]] + wc[b[u]] + r[u]), xr[u], yr[u], zr[u], (int)i);
        }
        __syncthreads();
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
        T xr, yr, zr;
        rescale_point<T, DIM>(x, y, z, i, g, xr, yr, zr, outside);  // rescale (+ slab clamp)
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

template <typename T, int DIM, int KM>
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
    // key table: counts -> exclusive offsets in entries [4, 4 + nkeys]; place_points advances them into end
    // offsets, after which entry 3 + k is the start of key k again (entry 3 stays 0): Plan::key_offsets()
    int *keybuf = p.keyoff.template as<int>();
    int *keyoff = keybuf + 4;
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
    CFB_CUDA_OK(cudaMemsetAsync(keybuf, 0, sizeof(int) * (size_t)(nscan + 4), st));
    // locality pass (see coarse_count_kernel): buckets of ~4 MB of records
    const PtRec<T> *part = nullptr;
    if (p.partitioned && M > 0) {
        static const long long bucket_env = [] { const char *e = getenv("CFB_SORT_BUCKET_MB"); return e ? atoll(e) << 20 : 0LL; }();   // experiments
        if (bucket_env > 0) p.sort_bucket_bytes = bucket_env;
        int nb = 32;
        while (nb < CP_MAXB && (long long)nb * p.sort_bucket_bytes < (long long)M * (long long)sizeof(PtRec<T>)) nb *= 2;
        const long long nrows = DIM == 1 ? p.nbin[0] : (long long)p.nbin[1] * (DIM > 2 ? p.nbin[2] : 1);   // rows of bins along x
        int kshift = 0;
        while (((nrows - 1) >> kshift) >= nb) ++kshift;
        nb = (int)((nrows - 1) >> kshift) + 1;
        int *ccnt = p.coarse.template as<int>();
        PtRec<T> *tmp = p.tmprecs.template as<PtRec<T>>();
        int nbits = 1;                                   // votes per point: the bits of a bucket number + the validity bit
        while ((1 << (nbits - 1)) < nb) ++nbits;
        const size_t csm = (size_t)(CP_WARPS + 2) * nb * sizeof(int);
        const long long tile = (long long)CP_THREADS * CP_PPT;
        const long long wt = (M + tile - 1) / tile;
        long long cb = wt < (long long)p.num_sms * CP_BPS ? wt : (long long)p.num_sms * CP_BPS;
        const long long tpb = (wt + cb - 1) / cb;          // tiles per block
        const int cblocks = (int)((wt + tpb - 1) / tpb);
        const long long nmat = (long long)nb * cblocks + 1;   // count matrix [bucket][block] + the trailing total
        const int mtiles = (int)((nmat + SCAN_TILE - 1) / SCAN_TILE);
        CFB_CUDA_OK(cudaMemsetAsync(ccnt + nmat - 1, 0, sizeof(int), st));
        coarse_count_kernel<T, DIM, KM><<<cblocks, CP_THREADS, csm, st>>>(M, x, y, z, g, kshift, nb, nbits, tpb, ccnt, p.slab ? scal + 3 : nullptr);
        scan_reduce_kernel<<<mtiles, SCAN_THREADS, 0, st>>>(nmat, ccnt, tilesum);
        scan_top_kernel<<<1, 1024, 0, st>>>(mtiles, tilesum);
        scan_apply_kernel<<<mtiles, SCAN_THREADS, 0, st>>>(nmat, ccnt, tilesum);
        coarse_scatter_kernel<T, DIM, KM><<<cblocks, CP_THREADS, csm, st>>>(M, x, y, z, g, kshift, nb, nbits, tpb, ccnt, tmp);
        p.launches_setpts += 5;
        part = tmp;
    }
    if (M > 0) {
        if (part) key_count_kernel<T, DIM, true, KM><<<blocks, threads, 0, st>>>(M, x, y, z, part, g, keyoff, nullptr);
        else key_count_kernel<T, DIM, false, KM><<<blocks, threads, 0, st>>>(M, x, y, z, nullptr, g, keyoff, p.slab ? scal + 3 : nullptr);
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
        if (part) place_points_kernel<T, DIM, true, KM><<<blocks, threads, 0, st>>>(M, x, y, z, part, g, keyoff, recs);
        else place_points_kernel<T, DIM, false, KM><<<blocks, threads, 0, st>>>(M, x, y, z, nullptr, g, keyoff, recs);
        p.launches_setpts++;
    }
    if (M > 0 && p.local_sort) {
        LocalSortArgs la;
        la.keyoff = p.key_offsets();
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
    for (int d = 0; d < 3; ++d) {
        auto lg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
        g.lg_ibs[d] = lg(g.ibs[d]); g.lg_spb[d] = lg(g.spb[d]);
        if (g.lg_spb[d] < 0) g.lg_ibs[d] = -1;
        g.inv_ibs[d] = g.lg_ibs[d] >= 0 ? 1.0f / (float)g.ibs[d] : 0.0f;
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
    // locality pass in front of the counting sort (coarse_count / coarse_scatter): it costs an extra read of the
    // coordinates and a read + write of the records, and pays when the one-level key table is far beyond L2 --
    // then every histogram reduction and every cursor atomic of the direct sort is a DRAM sector read-modify-write
    // (config 3: 537 MB table, setpts 8.2 -> 6.9 ms).  With an L2-resident table the direct scatter is faster
    // (measured: config 2 2.5 vs 3.1 ms, config 5's density 7.0 vs 10.7 ms; profiles/r02v).
    // cufinufft*_set_sort_levels: +4 never, +8 always.  It needs a second record array: when that does not
    // fit next to everything else the direct scatter still works.
    p.partitioned = p.sorted && p.M > 0 && p.sort_partition != 1 &&
                    (p.sort_partition == 2 || (one_level && (nkeys + 1) * 4 >= 2 * p.l2_bytes && M * sizeof(PtRec<T>) >= (size_t)(4 * p.l2_bytes)));
    if (p.partitioned) {
        if (p.tmprecs.reserve(M * sizeof(PtRec<T>)) != cudaSuccess || p.coarse.reserve(((size_t)CP_MAXB * p.num_sms * CP_BPS + 8) * sizeof(int)) != cudaSuccess) {
            cudaGetLastError();
            p.tmprecs.release();
            p.partitioned = false;
        }
    }
    if (p.sorted) {
        CFB_CUDA_OK(p.keyoff.reserve(((size_t)nkeys + 1 + 8) * sizeof(int)));
        CFB_CUDA_OK(p.tilesum.reserve((((size_t)(nkeys + 1) + (size_t)CP_MAXB * p.num_sms * CP_BPS) / SCAN_TILE + 4) * sizeof(int)));
    }
    size_t maxslots = (size_t)p.nbins + M / (size_t)p.opts.gpu_maxsubprobsize + 1;
    CFB_CUDA_OK(p.subprob_to_bin.reserve(maxslots * sizeof(int)));
    if (p.ilist) {
        CFB_CUDA_OK(p.isubstart.reserve(((size_t)p.nibins + 1 + 4) * sizeof(int)));
        CFB_CUDA_OK(p.is2b.reserve(((size_t)p.nibins + M / (size_t)p.imaxsub + 1) * sizeof(int)));
    }
    p.idx_valid = false;
    // compile-time specialisation of the key (key_any): power-of-two bin sizes in every dimension, and what the
    // key holds inside a bin
    int km = g.bankc > 0 ? 3 : (g.cpb == 1 ? 1 : 2);
    for (int d = 0; d < p.dim; ++d) if (g.lg_ibs[d] < 0) km = 0;
    if (p.key_generic) km = 0;
    switch (p.dim * 4 + km) {
        case 4: return setpts_dim<T, 1, 0>(p);
        case 5: return setpts_dim<T, 1, 1>(p);
        case 6: return setpts_dim<T, 1, 2>(p);
        case 7: return setpts_dim<T, 1, 3>(p);
        case 8: return setpts_dim<T, 2, 0>(p);
        case 9: return setpts_dim<T, 2, 1>(p);
        case 10: return setpts_dim<T, 2, 2>(p);
        case 11: return setpts_dim<T, 2, 3>(p);
        case 12: return setpts_dim<T, 3, 0>(p);
        case 13: return setpts_dim<T, 3, 1>(p);
        case 14: return setpts_dim<T, 3, 2>(p);
        default: return setpts_dim<T, 3, 3>(p);
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
