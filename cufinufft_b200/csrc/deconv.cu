// deconv.cu -- mode-space stages: phihat (Fourier series of the kernel), deconvolve
// (type-1 step 3) and amplify (type-2 step 1).
//
// Reference: FseriesKernelCompute src/common.cu:16-65; Deconvolve_{1,2,3}d /
// Amplify_{1,2,3}d and their per-transform launch loops src/deconvolve_wrapper.cu:14-254.
// Differences in schedule, not in arithmetic:
//  - one launch covers all transforms of the batch (blockIdx.y = transform);
//  - amplify writes the WHOLE fine grid (modes + zero padding) in one coalesced pass,
//    so the reference's cudaMemset(fw) before Amplify (deconvolve_wrapper.cu:141,181,228)
//    and its scattered writes disappear;
//  - x (fastest) index arithmetic is done per row segment, no per-element div/mod.
// Values: fk = fw / (phihat1*phihat2*phihat3) with the product then the division in T,
// exactly the reference's expression.
#include <algorithm>
#include <cuComplex.h>
#include "cfb_device.cuh"

namespace cfb {

// fwkerhalf_d[k] = sum_n f_n * 2 * |a_n|^k cos(k arg a_n), k = 0..nf_d/2; sum in double,
// accumulator narrowed to T after every term like the reference's `FLT x` (src/common.cu:39-43).
template <typename T>
__global__ void __launch_bounds__(128)
fseries_kernel(int nf1, int nf2, int nf3, int q, const T *__restrict__ f, const double *__restrict__ a_reim,
               T *__restrict__ k1, T *__restrict__ k2, T *__restrict__ k3)
{
    const int d = blockIdx.y;
    const int nf = d == 0 ? nf1 : (d == 1 ? nf2 : nf3);
    T *out = d == 0 ? k1 : (d == 1 ? k2 : k3);
    const T *fd = f + d * MAX_NQUAD;
    const double *ad = a_reim + 2 * d * MAX_NQUAD;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf / 2 + 1; i += gridDim.x * blockDim.x) {
        const int brk = (int)(0.5 + i);
        T x = (T)0.0;
        for (int n = 0; n < q; ++n) {
            const cuDoubleComplex an = make_cuDoubleComplex(ad[2 * n], ad[2 * n + 1]);
            const double mag = cuCabs(an), ang = atan2(an.y, an.x);
            x += fd[n] * 2 * (pow(mag, brk) * cos(brk * ang));
        }
        out[i] = x;
    }
}

template <typename T>
int stage_fseries(Plan<T> &p)
{
    T f[3 * MAX_NQUAD] = {0};
    double a[2 * 3 * MAX_NQUAD] = {0};
    const int nf3 = p.nf3_global();          // slab plans: phihat of the GLOBAL z grid
    const int nf[3] = {p.nf1, p.nf2, nf3};
    for (int d = 0; d < p.dim; ++d)
        fseries_precomp<T>(nf[d], p.ns, p.es_beta, p.es_c, p.es_halfwidth, f + d * MAX_NQUAD, a + 2 * d * MAX_NQUAD);
    struct Scoped : DevBuf { ~Scoped() { release(); } } bf, ba;      // freed on every return path
    CFB_CUDA_OK(bf.reserve(sizeof(f)));
    CFB_CUDA_OK(ba.reserve(sizeof(a)));
    T *d_f = bf.template as<T>();
    double *d_a = ba.template as<double>();
    CFB_CUDA_OK(cudaMemcpyAsync(d_f, f, sizeof(f), cudaMemcpyHostToDevice, p.stream));
    CFB_CUDA_OK(cudaMemcpyAsync(d_a, a, sizeof(a), cudaMemcpyHostToDevice, p.stream));
    const int q = (int)(2 + 3.0 * (T)(p.ns / 2.0));
    int nout = std::max(std::max(p.nf1, p.nf2), nf3) / 2 + 1;
    dim3 grid((nout + 127) / 128, p.dim);
    fseries_kernel<T><<<grid, 128, 0, p.stream>>>(p.nf1, p.nf2, nf3, q, d_f, d_a, p.fwker[0].template as<T>(),
                                                  p.fwker[1].template as<T>(), p.fwker[2].template as<T>());
    CFB_CUDA_OK(cudaGetLastError());
    CFB_CUDA_OK(cudaStreamSynchronize(p.stream));   // f/a are stack + temporaries: plan time only
    return 0;
}

// mode index i (0..m-1, k = i - m/2) -> fine-grid index
__device__ __forceinline__ int mode_to_grid(int i, int m, int nf)
{
    const int k = i - m / 2;
    return k >= 0 ? k : nf + k;
}

// Launch shape for both kernels: blockIdx.x = row ((y,z) flattened, up to 2^31-1 rows),
// blockIdx.y = chunk of 1024 consecutive x entries, blockIdx.z = transform.

// type 1: fk[t][k3][k2][k1] = fw[t][w3][w2][w1] / (phihat product).  x fastest: coalesced
// 8/16-byte stores; loads coalesced on the two x segments (modes >= 0, modes < 0).
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
deconvolve_kernel(int ms, int mt, int mu, int nf1, int nf2, int nf3, const typename cplx_of<T>::type *__restrict__ fw,
                  typename cplx_of<T>::type *__restrict__ fk, const T *__restrict__ ker1, const T *__restrict__ ker2,
                  const T *__restrict__ ker3, long long fwstride, long long fkstride)
{
    using C = typename cplx_of<T>::type;
    const int row = blockIdx.x;                 // (k2, k3) flattened
    const int t = blockIdx.z;
    const int k2 = DIM > 1 ? row % mt : 0, k3 = DIM > 2 ? row / mt : 0;
    T ky = 1, kz = 1;
    size_t in_row = 0;
    if (DIM > 1) { ky = ker2[abs(k2 - mt / 2)]; in_row += (size_t)mode_to_grid(k2, mt, nf2) * nf1; }
    if (DIM > 2) { kz = ker3[abs(k3 - mu / 2)]; in_row += (size_t)mode_to_grid(k3, mu, nf3) * nf1 * nf2; }
    const C *src = fw + (size_t)t * fwstride + in_row;
    C *dst = fk + (size_t)t * fkstride + (size_t)row * ms;
    const int x0 = blockIdx.y * 1024;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k1 = x0 + j * 256 + threadIdx.x;
        if (k1 < ms) {
            T kv = ker1[abs(k1 - ms / 2)];
            if (DIM > 1) kv = kv * ky;
            if (DIM > 2) kv = kv * kz;
            const C v = src[mode_to_grid(k1, ms, nf1)];
            C o; o.x = v.x / kv; o.y = v.y / kv;
            dst[k1] = o;
        }
    }
}

// fine-grid index -> mode index (0..m-1) or -1 when the cell is zero padding
__device__ __forceinline__ int grid_to_mode(int w, int m, int nf)
{
    if (w <= (m - 1) / 2) return w + m / 2;
    if (w >= nf - m / 2) return w - nf + m / 2;
    return -1;
}

// type 2: fw[t][w3][w2][w1] = in-range mode ? fk[...]/(phihat product) : 0, whole grid.
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
amplify_kernel(int ms, int mt, int mu, int nf1, int nf2, int nf3, typename cplx_of<T>::type *__restrict__ fw,
               const typename cplx_of<T>::type *__restrict__ fk, const T *__restrict__ ker1, const T *__restrict__ ker2,
               const T *__restrict__ ker3, long long fwstride, long long fkstride)
{
    using C = typename cplx_of<T>::type;
    const int row = blockIdx.x;                 // (w2, w3) flattened
    const int t = blockIdx.z;
    const int w2 = DIM > 1 ? row % nf2 : 0, w3 = DIM > 2 ? row / nf2 : 0;
    const int i2 = DIM > 1 ? grid_to_mode(w2, mt, nf2) : 0;
    const int i3 = DIM > 2 ? grid_to_mode(w3, mu, nf3) : 0;
    const bool row_live = i2 >= 0 && i3 >= 0;
    T ky = 1, kz = 1;
    if (row_live && DIM > 1) ky = ker2[abs(i2 - mt / 2)];
    if (row_live && DIM > 2) kz = ker3[abs(i3 - mu / 2)];
    C *dst = fw + (size_t)t * fwstride + (size_t)row * nf1;
    const C *src = fk + (size_t)t * fkstride + ((size_t)(row_live ? i3 : 0) * mt + (row_live ? i2 : 0)) * ms;
    const int x0 = blockIdx.y * 1024;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int w1 = x0 + j * 256 + threadIdx.x;
        if (w1 < nf1) {
            C o; o.x = 0; o.y = 0;
            const int i1 = grid_to_mode(w1, ms, nf1);
            if (row_live && i1 >= 0) {
                T kv = ker1[abs(i1 - ms / 2)];
                if (DIM > 1) kv = kv * ky;
                if (DIM > 2) kv = kv * kz;
                const C v = src[i1];
                o.x = v.x / kv; o.y = v.y / kv;
            }
            dst[w1] = o;
        }
    }
}

template <typename T>
int stage_deconvolve(Plan<T> &p, typename Plan<T>::C *fk, const typename Plan<T>::C *fw, int nt)
{
    const T *k1 = p.fwker[0].template as<T>(), *k2 = p.fwker[1].template as<T>(), *k3 = p.fwker[2].template as<T>();
    dim3 grid(p.mt * p.mu, (p.ms + 1023) / 1024, nt);
    const long long fws = (long long)p.grid_cells(), fks = (long long)p.nmodes();
    switch (p.dim) {
        case 1: deconvolve_kernel<T, 1><<<grid, 256, 0, p.stream>>>(p.ms, p.mt, p.mu, p.nf1, p.nf2, p.nf3, fw, fk, k1, k2, k3, fws, fks); break;
        case 2: deconvolve_kernel<T, 2><<<grid, 256, 0, p.stream>>>(p.ms, p.mt, p.mu, p.nf1, p.nf2, p.nf3, fw, fk, k1, k2, k3, fws, fks); break;
        default: deconvolve_kernel<T, 3><<<grid, 256, 0, p.stream>>>(p.ms, p.mt, p.mu, p.nf1, p.nf2, p.nf3, fw, fk, k1, k2, k3, fws, fks); break;
    }
    p.launches_exec++;
    CFB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T>
int stage_amplify(Plan<T> &p, const typename Plan<T>::C *fk, typename Plan<T>::C *fw, int nt)
{
    const T *k1 = p.fwker[0].template as<T>(), *k2 = p.fwker[1].template as<T>(), *k3 = p.fwker[2].template as<T>();
    dim3 grid(p.nf2 * p.nf3, (p.nf1 + 1023) / 1024, nt);
    const long long fws = (long long)p.grid_cells(), fks = (long long)p.nmodes();
    switch (p.dim) {
        case 1: amplify_kernel<T, 1><<<grid, 256, 0, p.stream>>>(p.ms, p.mt, p.mu, p.nf1, p.nf2, p.nf3, fw, fk, k1, k2, k3, fws, fks); break;
        case 2: amplify_kernel<T, 2><<<grid, 256, 0, p.stream>>>(p.ms, p.mt, p.mu, p.nf1, p.nf2, p.nf3, fw, fk, k1, k2, k3, fws, fks); break;
        default: amplify_kernel<T, 3><<<grid, 256, 0, p.stream>>>(p.ms, p.mt, p.mu, p.nf1, p.nf2, p.nf3, fw, fk, k1, k2, k3, fws, fks); break;
    }
    p.launches_exec++;
    CFB_CUDA_OK(cudaGetLastError());
    return 0;
}

template int stage_fseries<float>(Plan<float> &);
template int stage_fseries<double>(Plan<double> &);
template int stage_deconvolve<float>(Plan<float> &, float2 *, const float2 *, int);
template int stage_deconvolve<double>(Plan<double> &, double2 *, const double2 *, int);
template int stage_amplify<float>(Plan<float> &, const float2 *, float2 *, int);
template int stage_amplify<double>(Plan<double> &, const double2 *, double2 *, int);

}  // namespace cfb
