// slab.cu -- one rank's share of a SINGLE large 3-D transform split into z-slabs of the fine grid
// (BASELINE.json config 5).
//
// The reference has no multi-GPU path for one transform (a plan lives on one device,
// src/cufinufft.cu:101-110); what is partitioned here is its 3-D pipeline
// src/3d/cufinufft3d.cu:15-165: spread|interp on fw[nf3][nf2][nf1], a 3-D cuFFT, deconvolve|amplify.
//
// Rank r of G owns the fine-grid planes [z0, z1) and holds them with pad = ceil(ns/2) halo planes
// on both sides (local plane l = global plane z0 - pad + l, periodic), and owns the points whose
// stencils lie inside (floor(z_r) in [z0, z1)).  The 3-D FFT is separated into its z part on
// MODE columns and its (x,y) part on PLANES, and the small mode array is replicated instead of
// transposing the big grid:
//   type 2:  fk (replicated) --amplify_z--> zbuf[nf3g][mt][ms] (all three phihat factors applied,
//            zero padded in z) --1-D cuFFT along z (strided, ms*mt columns: 1/4 of the fine grid's
//            columns)--> planes;  the nz+2*pad planes this rank needs -- halo included, so NO
//            exchange at all -- are zero-padded in (x,y) into the local grid (scatter_xy), 2-D
//            cuFFT per plane, interpolation at the rank's points.        Collectives: none.
//   type 1:  spread into the local haloed grid; the 2*pad halo planes are ADDED into the
//            z-neighbours' edge planes (halo_pack -> NCCL send/recv by the caller -> halo_add);
//            2-D cuFFT of the owned planes; the ms*mt mode columns of these planes go into zbuf
//            (zero elsewhere), 1-D cuFFT along z, deconvolve: a PARTIAL fk that the caller sums
//            over ranks (all-reduce of the small mode array).   Collectives: halo add, all-reduce.
// The 17 GB fine grid of config 5 never crosses NVLink; the price is a redundant z-FFT of the
// 4.3 GB zbuf on every rank and 2*pad/nz extra planes in the 2-D FFT (7.8 % at G = 8).
// Kernel weights are evaluated from the GLOBAL rescaled coordinate (records keep it, SIArgs::zshift
// moves the stencil start), so spread/interp arithmetic is bit-identical to the undivided plan.
#include <cuComplex.h>
#include <type_traits>
#include "cfb_device.cuh"
#include "../../include/cufinufft_b200.h"

namespace cfb {

__device__ __forceinline__ int slab_mode_to_grid(int i, int m, int nf)
{
    const int k = i - m / 2;
    return k >= 0 ? k : nf + k;
}
__device__ __forceinline__ int slab_grid_to_mode(int w, int m, int nf)
{
    if (w <= (m - 1) / 2) return w + m / 2;
    if (w >= nf - m / 2) return w - nf + m / 2;
    return -1;
}

// Launch shape of the row kernels: blockIdx.x = row, blockIdx.y = chunk of 1024 x entries.

// type 2, step 1: zbuf[w3][i2][i1] = fk[i3(w3)][i2][i1] / (phihat1*phihat2*phihat3), zero where w3
// is padding.  Same value expression as Amplify_3d (src/deconvolve_wrapper.cu:98-121).
template <typename T>
__global__ void __launch_bounds__(256)
slab_amplify_z_kernel(int ms, int mt, int mu, int nf3g, typename cplx_of<T>::type *__restrict__ zbuf,
                      const typename cplx_of<T>::type *__restrict__ fk, const T *__restrict__ ker1,
                      const T *__restrict__ ker2, const T *__restrict__ ker3)
{
    using C = typename cplx_of<T>::type;
    const int row = blockIdx.x;                       // (i2, w3) flattened, i2 fastest
    const int i2 = row % mt, w3 = row / mt;
    const int i3 = slab_grid_to_mode(w3, mu, nf3g);
    C *dst = zbuf + (size_t)row * ms;
    const C *src = fk + ((size_t)(i3 >= 0 ? i3 : 0) * mt + i2) * ms;
    T ky = 1, kz = 1;
    if (i3 >= 0) { ky = ker2[abs(i2 - mt / 2)]; kz = ker3[abs(i3 - mu / 2)]; }
    const int x0 = blockIdx.y * 1024;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i1 = x0 + j * 256 + threadIdx.x;
        if (i1 < ms) {
            C o; o.x = 0; o.y = 0;
            if (i3 >= 0) {
                T kv = ker1[abs(i1 - ms / 2)];
                kv = kv * ky;
                kv = kv * kz;
                const C v = src[i1];
                o.x = v.x / kv; o.y = v.y / kv;
            }
            dst[i1] = o;
        }
    }
}

// type 2, step 3: local plane l (global plane wrap(zshift + l)) of the fine grid =
// the mode columns of zbuf zero-padded in (x,y).  Writes the WHOLE local grid.
template <typename T>
__global__ void __launch_bounds__(256)
slab_scatter_xy_kernel(int ms, int mt, int nf1, int nf2, int nf3g, int zshift, typename cplx_of<T>::type *__restrict__ fw,
                       const typename cplx_of<T>::type *__restrict__ zbuf)
{
    using C = typename cplx_of<T>::type;
    const int row = blockIdx.x;                       // (w2, l) flattened, w2 fastest
    const int w2 = row % nf2, l = row / nf2;
    int g = zshift + l;
    g = g < 0 ? g + nf3g : (g >= nf3g ? g - nf3g : g);
    const int i2 = slab_grid_to_mode(w2, mt, nf2);
    C *dst = fw + (size_t)row * nf1;
    const C *src = zbuf + ((size_t)g * mt + (i2 >= 0 ? i2 : 0)) * ms;
    const int x0 = blockIdx.y * 1024;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int w1 = x0 + j * 256 + threadIdx.x;
        if (w1 < nf1) {
            C o; o.x = 0; o.y = 0;
            const int i1 = slab_grid_to_mode(w1, ms, nf1);
            if (i2 >= 0 && i1 >= 0) o = src[i1];
            dst[w1] = o;
        }
    }
}

// type 1, step 4: mode columns of the owned planes -> zbuf[g][i2][i1] (the other planes of zbuf
// were zeroed): the (x,y) truncation of Deconvolve_3d, phihat division deferred to the z stage.
template <typename T>
__global__ void __launch_bounds__(256)
slab_gather_xy_kernel(int ms, int mt, int nf1, int nf2, int z0, int pad, const typename cplx_of<T>::type *__restrict__ fw,
                      typename cplx_of<T>::type *__restrict__ zbuf)
{
    using C = typename cplx_of<T>::type;
    const int row = blockIdx.x;                       // (i2, owned plane) flattened, i2 fastest
    const int i2 = row % mt, lz = row / mt;
    const C *src = fw + ((size_t)(pad + lz) * nf2 + slab_mode_to_grid(i2, mt, nf2)) * nf1;
    C *dst = zbuf + ((size_t)(z0 + lz) * mt + i2) * ms;
    const int x0 = blockIdx.y * 1024;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i1 = x0 + j * 256 + threadIdx.x;
        if (i1 < ms) dst[i1] = src[slab_mode_to_grid(i1, ms, nf1)];
    }
}

// type 1, step 6: fk_partial[i3][i2][i1] = zbuf[w3(i3)][i2][i1] / (phihat1*phihat2*phihat3)
// (Deconvolve_3d's expression, src/deconvolve_wrapper.cu:52-75).
template <typename T>
__global__ void __launch_bounds__(256)
slab_deconvolve_z_kernel(int ms, int mt, int mu, int nf3g, const typename cplx_of<T>::type *__restrict__ zbuf,
                         typename cplx_of<T>::type *__restrict__ fk, const T *__restrict__ ker1,
                         const T *__restrict__ ker2, const T *__restrict__ ker3)
{
    using C = typename cplx_of<T>::type;
    const int row = blockIdx.x;                       // (i2, i3) flattened
    const int i2 = row % mt, i3 = row / mt;
    const T ky = ker2[abs(i2 - mt / 2)], kz = ker3[abs(i3 - mu / 2)];
    const C *src = zbuf + ((size_t)slab_mode_to_grid(i3, mu, nf3g) * mt + i2) * ms;
    C *dst = fk + (size_t)row * ms;
    const int x0 = blockIdx.y * 1024;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i1 = x0 + j * 256 + threadIdx.x;
        if (i1 < ms) {
            T kv = ker1[abs(i1 - ms / 2)];
            kv = kv * ky;
            kv = kv * kz;
            const C v = src[i1];
            C o; o.x = v.x / kv; o.y = v.y / kv;
            dst[i1] = o;
        }
    }
}

// dst[i] += src[i] over n reals, 16-byte vectors (halo planes are contiguous and 16-byte aligned)
template <typename T>
__global__ void __launch_bounds__(256)
slab_add_kernel(size_t nvec, T *__restrict__ dst, const T *__restrict__ src)
{
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    V *d = reinterpret_cast<V *>(dst);
    const V *s = reinterpret_cast<const V *>(src);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        V a = d[i];
        const V b = s[i];
        if constexpr (sizeof(T) == 4) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
        else { a.x += b.x; a.y += b.y; }
        d[i] = a;
    }
}

static cufftResult slab_fft(cufftHandle h, float2 *d, int dir) { return cufftExecC2C(h, d, d, dir); }
static cufftResult slab_fft(cufftHandle h, double2 *d, int dir) { return cufftExecZ2Z(h, d, d, dir); }

template <typename T>
int slab_make_ffts(Plan<T> &p)
{
    const cufftType ty = sizeof(T) == 4 ? CUFFT_C2C : CUFFT_Z2Z;
    int n2[2] = {p.nf2, p.nf1};
    const int planes = p.type == 2 ? p.nf3 : p.z1 - p.z0;        // type 2 transforms the halo planes too
    if (cufftPlanMany(&p.fft2d, 2, n2, n2, 1, p.nf1 * p.nf2, n2, 1, p.nf1 * p.nf2, ty, planes) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
    p.have_fft2d = true;
    int nz[1] = {p.nf3g};
    const int cols = p.ms * p.mt;
    if (cufftPlanMany(&p.fftz, 1, nz, nz, cols, 1, nz, cols, 1, ty, cols) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
    p.have_fftz = true;
    return 0;
}

template <typename T>
int slab_type2(Plan<T> &p, typename Plan<T>::C *c, const typename Plan<T>::C *fk)
{
    using C = typename Plan<T>::C;
    cudaStream_t st = p.stream;
    p.launches_exec = 0;
    if (p.timing) for (auto &e : p.ev) if (!e) cudaEventCreate(&e);
    auto mark = [&](int i) { if (p.timing) cudaEventRecord(p.ev[i], st); };
    const T *k1 = p.fwker[0].template as<T>(), *k2 = p.fwker[1].template as<T>(), *k3 = p.fwker[2].template as<T>();
    C *zbuf = p.zbuf.template as<C>(), *fw = p.fw.template as<C>();
    if (cufftSetStream(p.fftz, st) != CUFFT_SUCCESS || cufftSetStream(p.fft2d, st) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
    mark(0); mark(1);
    slab_amplify_z_kernel<T><<<dim3(p.mt * p.nf3g, (p.ms + 1023) / 1024), 256, 0, st>>>(p.ms, p.mt, p.mu, p.nf3g, zbuf, fk, k1, k2, k3);
    if (slab_fft(p.fftz, zbuf, p.iflag) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
    slab_scatter_xy_kernel<T><<<dim3(p.nf2 * p.nf3, (p.nf1 + 1023) / 1024), 256, 0, st>>>(p.ms, p.mt, p.nf1, p.nf2, p.nf3g, p.zshift, fw, zbuf);
    p.launches_exec += 2;
    mark(2);
    if (slab_fft(p.fft2d, fw, p.iflag) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
    mark(3);
    if (int e = stage_interp(p, c, fw, 1)) return e;
    mark(4);
    CFB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T>
int slab_type1_spread(Plan<T> &p, const typename Plan<T>::C *c)
{
    using C = typename Plan<T>::C;
    p.launches_exec = 0;
    if (p.timing) for (auto &e : p.ev) if (!e) cudaEventCreate(&e);
    if (p.timing) cudaEventRecord(p.ev[0], p.stream);
    CFB_CUDA_OK(cudaMemsetAsync(p.fw.p, 0, p.grid_cells() * sizeof(C), p.stream));
    if (p.timing) cudaEventRecord(p.ev[1], p.stream);
    if (int e = stage_spread(p, c, p.fw.template as<C>(), 1)) return e;
    if (p.timing) cudaEventRecord(p.ev[2], p.stream);
    return 0;
}

// side 0 = the low halo planes [0, pad), side 1 = the high halo planes [pad + nz, nz + 2 pad)
template <typename T>
int slab_halo_pack(Plan<T> &p, int side, typename Plan<T>::C *buf)
{
    using C = typename Plan<T>::C;
    const size_t plane = (size_t)p.nf1 * p.nf2, pad = (size_t)p.tile_pad;
    const C *src = p.fw.template as<C>() + (side == 0 ? 0 : (pad + (size_t)(p.z1 - p.z0)) * plane);
    CFB_CUDA_OK(cudaMemcpyAsync(buf, src, pad * plane * sizeof(C), cudaMemcpyDeviceToDevice, p.stream));
    return 0;
}

// side 0: `buf` = the PREVIOUS rank's high halo -> added into this rank's first owned planes
// [pad, 2 pad); side 1: the NEXT rank's low halo -> added into the last owned planes [nz, nz + pad).
template <typename T>
int slab_halo_add(Plan<T> &p, int side, const typename Plan<T>::C *buf)
{
    using C = typename Plan<T>::C;
    const size_t plane = (size_t)p.nf1 * p.nf2, pad = (size_t)p.tile_pad;
    C *dst = p.fw.template as<C>() + (side == 0 ? pad : (size_t)(p.z1 - p.z0)) * plane;
    const size_t nvec = pad * plane * sizeof(C) / 16;
    const size_t want = (nvec + 255) / 256;
    const int blocks = (int)(want > (size_t)p.num_sms * 16 ? (size_t)p.num_sms * 16 : (want < 1 ? 1 : want));
    slab_add_kernel<T><<<blocks, 256, 0, p.stream>>>(nvec, reinterpret_cast<T *>(dst), reinterpret_cast<const T *>(buf));
    p.launches_exec++;
    CFB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T>
int slab_type1_finish(Plan<T> &p, typename Plan<T>::C *fk_partial)
{
    using C = typename Plan<T>::C;
    cudaStream_t st = p.stream;
    const T *k1 = p.fwker[0].template as<T>(), *k2 = p.fwker[1].template as<T>(), *k3 = p.fwker[2].template as<T>();
    C *zbuf = p.zbuf.template as<C>(), *fw = p.fw.template as<C>();
    const int nz = p.z1 - p.z0;
    if (cufftSetStream(p.fftz, st) != CUFFT_SUCCESS || cufftSetStream(p.fft2d, st) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
    if (slab_fft(p.fft2d, fw + (size_t)p.tile_pad * p.nf1 * p.nf2, p.iflag) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
    if (p.timing) cudaEventRecord(p.ev[3], st);
    CFB_CUDA_OK(cudaMemsetAsync(zbuf, 0, (size_t)p.nf3g * p.mt * p.ms * sizeof(C), st));
    slab_gather_xy_kernel<T><<<dim3(p.mt * nz, (p.ms + 1023) / 1024), 256, 0, st>>>(p.ms, p.mt, p.nf1, p.nf2, p.z0, p.tile_pad, fw, zbuf);
    if (slab_fft(p.fftz, zbuf, p.iflag) != CUFFT_SUCCESS) return CFB_ERR_CUFFT;
    slab_deconvolve_z_kernel<T><<<dim3(p.mt * p.mu, (p.ms + 1023) / 1024), 256, 0, st>>>(p.ms, p.mt, p.mu, p.nf3g, zbuf, fk_partial, k1, k2, k3);
    p.launches_exec += 2;
    if (p.timing) cudaEventRecord(p.ev[4], st);
    CFB_CUDA_OK(cudaGetLastError());
    return 0;
}

#define CFB_INST_SLAB(T)                                                                   \
    template int slab_make_ffts<T>(Plan<T> &);                                             \
    template int slab_type2<T>(Plan<T> &, typename Plan<T>::C *, const typename Plan<T>::C *); \
    template int slab_type1_spread<T>(Plan<T> &, const typename Plan<T>::C *);             \
    template int slab_halo_pack<T>(Plan<T> &, int, typename Plan<T>::C *);                 \
    template int slab_halo_add<T>(Plan<T> &, int, const typename Plan<T>::C *);            \
    template int slab_type1_finish<T>(Plan<T> &, typename Plan<T>::C *);
CFB_INST_SLAB(float)
CFB_INST_SLAB(double)

}  // namespace cfb
