// cfb_device.cuh -- device-side building blocks shared by the setpts / spread /
// interp kernels: coordinate rescale, bin index, ES-kernel evaluation.
#pragma once
#include <cuda_runtime.h>
#include "cfb_plan.h"

namespace cfb {

// ---- RESCALE (pirange=1): reference contrib/spreadinterp.h:36-38 -------------
// (x/2pi + {0.5|1.5|-0.5}) * nf, evaluated in DOUBLE for both precisions (the
// reference's literals are double) with the multiply-add contracted as nvcc's
// default -fmad=true does for that expression, then narrowed to T.
template <typename T>
__device__ __forceinline__ T rescale(T x, int nf)
{
    const T pi = (T)3.14159265358979323846;
    double shift = (x < -pi) ? 1.5 : ((x >= pi) ? -0.5 : 0.5);
    return (T)(fma((double)x, 0.159154943091895336, shift) * (double)nf);
}

// bin coordinate with the reference's two clamps (src/2d/spreadinterp2d.cu:113-118)
template <typename T>
__device__ __forceinline__ int bin_coord(T xr, int binsize, int nbin)
{
    int b = (int)floor(xr / (T)binsize);
    b = b >= nbin ? b - 1 : b;
    b = b < 0 ? 0 : b;
    return b;
}

// first grid index touched by a point: ceil(xr - ns/2) in double (src/2d/spreadinterp2d.cu:35)
template <typename T>
__device__ __forceinline__ int stencil_start(T xr, int ns)
{
    return (int)ceil((double)xr - ns * 0.5);
}

// Stencil-origin cell of a point inside its bin: (stencil_start - first possible start of the
// bin), clamped to [0, nk).  Only a SORT KEY (points of one key share their stencil origin, up
// to the clamped edge cases); results never depend on it.
template <typename T>
__device__ __forceinline__ int stencil_cell(T xr, int ns, int bin_origin, int nk)
{
    int k = stencil_start(xr, ns) - (bin_origin - ns / 2);
    if (!(ns & 1)) k -= 1;          // even widths: origin changes at integers, k = 0 only for xr exactly on the bin edge
    return k < 0 ? 0 : (k >= nk ? nk - 1 : k);
}

__device__ __forceinline__ int rec_index(const PtRec<float> &r) { return r.idx; }
__device__ __forceinline__ int rec_index(const PtRec<double> &r) { return (int)r.idx; }

__device__ __forceinline__ PtRec<float> load_rec(const PtRec<float> *p)
{
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    PtRec<float> r; r.x = v.x; r.y = v.y; r.z = v.z; r.idx = __float_as_int(v.w);
    return r;
}
__device__ __forceinline__ PtRec<double> load_rec(const PtRec<double> *p)
{
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    PtRec<double> r; r.x = a.x; r.y = a.y; r.z = b.x; r.idx = __double_as_longlong(b.y);
    return r;
}

// ---- exp(beta*sqrt(1 - c x^2)) for |x| < ns/2, un-normalised as in the reference
// (src/cuspreadinterp.h:6-16).  fp64 plans evaluate in double (bit-compatible formula);
// fp32 plans evaluate in fp32 (the reference promotes to double here; our deviation is
// <= ~1e-6 relative, budgeted inside the 1e-5 parity tolerance).
__device__ __forceinline__ float es_eval(float ax, float es_c, float es_beta, float half)
{
    float t = fmaf(-es_c * ax, ax, 1.0f);
    float v = expf(es_beta * sqrtf(fmaxf(t, 0.0f)));
    return ax < half ? v : 0.0f;
}
__device__ __forceinline__ double es_eval(double ax, double es_c, double es_beta, double half)
{
    double t = 1.0 - es_c * ax * ax;
    double v = exp(es_beta * sqrt(fmax(t, 0.0)));
    return ax < half ? v : 0.0;
}

// Kernel vector of one coordinate: ker[i] = phi(|x1 + i|), i < NS  (eval_kernel_vec,
// src/cuspreadinterp.h:33-40) or the Horner piecewise polynomial (:18-31) with
// coefficients hc[k*NS + i] (T, staged in shared memory by the caller).
template <typename T, int NS>
__device__ __forceinline__ void kernel_vector(T *ker, T x1, T es_c, T es_beta, bool horner, const T *hc, int ncoef)
{
    if (!horner) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            T a = x1 + (T)i;
            a = a < 0 ? -a : a;
            ker[i] = es_eval(a, es_c, es_beta, (T)(NS * 0.5));
        }
    } else {
        T z = (T)(2 * (double)x1 + NS - 1.0);
#pragma unroll
        for (int i = 0; i < NS; ++i) ker[i] = hc[(ncoef - 1) * NS + i];
        for (int k = ncoef - 2; k >= 0; --k) {
#pragma unroll
            for (int i = 0; i < NS; ++i) ker[i] = fma(z, ker[i], hc[k * NS + i]);
        }
    }
}

__device__ __forceinline__ int wrap_index(int i, int nf) { return i < 0 ? i + nf : (i >= nf ? i - nf : i); }

// vector reduction into global memory: one RED per complex cell
__device__ __forceinline__ void red_add(float2 *addr, float re, float im)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(re), "f"(im) : "memory");
}
__device__ __forceinline__ void red_add(double2 *addr, double re, double im)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(&addr->x), "d"(re) : "memory");
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(&addr->y), "d"(im) : "memory");
}

}  // namespace cfb
