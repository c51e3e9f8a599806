// cfb_device.cuh -- device-side building blocks shared by the setpts / spread /
// interp kernels: coordinate rescale, bin index, ES-kernel evaluation.
#pragma once
#include <cuda_runtime.h>
#include "cfb_plan.h"
#include "horner_const.cuh"

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
// fp32: the same integer without the double-precision convert (F2I.F64.CEIL took 18 % of the stall
// samples of the config-1 spread kernel, profiles/r01p): for even ns ceil(xr - ns/2) = ceil(xr) - ns/2,
// for odd ns ceil(xr - 1/2) = floor(xr) + (frac(xr) > 1/2); floor, the difference and the compare are
// exact in fp32, so the result is the exact ceiling; the double evaluation agrees for every x_r >= 2^-50
// (below that IT rounds x_r - ns/2 to an integer; such a point only moves one support-edge weight).
__device__ __forceinline__ int stencil_start(double xr, int ns)
{
    return (int)ceil(xr - ns * 0.5);
}
__device__ __forceinline__ int stencil_start(float xr, int ns)
{
    if (!(ns & 1)) return (int)ceilf(xr) - (ns >> 1);
    const float fl = floorf(xr);
    return (int)fl + ((xr - fl) > 0.5f ? 1 : 0) - ((ns - 1) >> 1);
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
// (src/cuspreadinterp.h:6-16, which evaluates sqrt and exp in double even in the fp32 build).
// Both versions are straight-line code without the special-case branches of libdevice:
//   sqrt(t), t in (0,1]: MUFU rsqrt seed + Newton steps;  exp(y), y in [0, 40]: y = n ln2 + r,
//   2^n applied to the exponent field.
// fp32 (14 instructions, 2 MUFU): s = sqrt(t) by one Newton step; the exponent in base 2 is formed as
//   s * (beta log2 e) with the constant split in two floats (no rounding of beta*s), n by the
//   1.5*2^23 rounding trick, 2^f by ex2.approx, 2^n added into the exponent field with one integer op.
//   Relative error ~1e-7 per weight plus beta * (the 6e-8 rounding of t) / (2 s).   fp64: degree-13
//   polynomial, ~1 ulp.
// The argument may have either sign.  |x| >= ns/2 (where the reference returns 0) is the CALLER's case:
// inside a stencil it can only be the first point, kernel_vector handles it.
struct EsConst32 { float c, bl_hi, bl_lo; };
__device__ __forceinline__ EsConst32 es_const(float es_c, float es_beta)
{
    const float L = 1.4426950216293335f, Llo = 1.9259629911266175e-08f;     // log2(e) = L + Llo
    EsConst32 k;
    k.c = es_c;
    k.bl_hi = es_beta * L;
    k.bl_lo = fmaf(es_beta, Llo, fmaf(es_beta, L, -k.bl_hi));
    return k;
}
__device__ __forceinline__ float es_eval32(float a, const EsConst32 &k)
{
    float t = fmaf(-k.c * a, a, 1.0f);
    t = fmaxf(t, 1e-30f);                    // support edge: keep rsqrt finite (the weight there is exp(0) either way)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    float s = t * r;
    const float e = fmaf(-s, s, t);
    s = fmaf(e * 0.5f, r, s);
    const float magic = 12582912.0f;         // 1.5 * 2^23: z's low mantissa bits hold n = rint(s * beta log2 e)
    const float z = fmaf(s, k.bl_hi, magic);
    const float n = z - magic;
    float f = fmaf(s, k.bl_hi, -n);
    f = fmaf(s, k.bl_lo, f);
    float p;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(f));
    return __int_as_float(__float_as_int(p) + (__float_as_int(z) << 23));      // p * 2^n (the magic's own bits shift out)
}
__device__ __forceinline__ float es_eval(float ax, float es_c, float es_beta, float half)
{
    const float v = es_eval32(ax, es_const(es_c, es_beta));
    return ax < half ? v : 0.0f;
}
__device__ __forceinline__ double es_eval(double ax, double es_c, double es_beta, double half)
{
    const double t = 1.0 - es_c * ax * ax;
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(t));
    double s = t * r;
    const double h = 0.5 * r;
    double e = fma(-s, s, t);
    s = fma(e, h, s);
    e = fma(-s, s, t);
    s = fma(e, h, s);
    e = fma(-s, s, t);
    s = fma(e, h, s);
    s = t > 0.0 ? s : 0.0;
    const double y = es_beta * s;
    const double magic = 6755399441055744.0;                                // 1.5 * 2^52
    const double z = fma(y, 1.4426950408889634, magic);
    const int n = __double2loint(z);
    const double nd = z - magic;
    double q = fma(-nd, 6.93147180369123816490e-01, y);                     // ln2 high 32 bits (exact product)
    q = fma(-nd, 1.90821492927058770002e-10, q);
    double p = 1.6059043836821613e-10;                                      // 1/13!
    p = fma(p, q, 2.08767569878681e-09);
    p = fma(p, q, 2.505210838544172e-08);
    p = fma(p, q, 2.755731922398589e-07);
    p = fma(p, q, 2.7557319223985893e-06);
    p = fma(p, q, 2.48015873015873e-05);
    p = fma(p, q, 1.984126984126984e-04);
    p = fma(p, q, 1.388888888888889e-03);
    p = fma(p, q, 8.333333333333333e-03);
    p = fma(p, q, 4.1666666666666664e-02);
    p = fma(p, q, 1.6666666666666666e-01);
    p = fma(p, q, 0.5);
    p = fma(p, q, 1.0);
    p = fma(p, q, 1.0);
    p = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return ax < half ? p : 0.0;
}

// Kernel values of one coordinate: dst[slot] = phi(|x1 + i|) for the NS stencil points
// (eval_kernel_vec, src/cuspreadinterp.h:33-40) or the Horner piecewise polynomial (:18-31) with
// coefficients hc[k*NS + i] (T, staged in shared memory by the caller).  `dst` may be a register
// array (UNROLL) or shared memory.
template <typename T, int NS, bool UNROLL>
__device__ __forceinline__ void kernel_vector(T *dst, T x1, T es_c, T es_beta, bool horner, const T *hc, int ncoef)
{
    if (!horner) {
        if constexpr (sizeof(T) == 4) {
            // x1 = xstart - x_r lies in [-ns/2, -ns/2 + 1): |x1 + i| < ns/2 for every i >= 1, and for i = 0
            // unless x1 == -ns/2 exactly (then the reference's abs(x) < ns/2 test gives 0)
            const EsConst32 k = es_const(es_c, es_beta);
            if (UNROLL) {
#pragma unroll
                for (int i = 0; i < NS; ++i) dst[i] = es_eval32(x1 + (T)i, k);
            } else {
#pragma unroll 1
                for (int i = 0; i < NS; ++i) dst[i] = es_eval32(x1 + (T)i, k);
            }
            if (!(x1 > (T)(-0.5 * NS))) dst[0] = 0;
        } else if (UNROLL) {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                T a = x1 + (T)i;
                a = a < 0 ? -a : a;
                dst[i] = es_eval(a, es_c, es_beta, (T)(NS * 0.5));
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < NS; ++i) {
                T a = x1 + (T)i;
                a = a < 0 ? -a : a;
                dst[i] = es_eval(a, es_c, es_beta, (T)(NS * 0.5));
            }
        }
    } else {
        // Horner: coefficients straight from constant memory (horner_const.cuh; every index is a compile-time
        // constant once unrolled: the coefficient is the constant-bank operand of its FMA).  `hc` / `ncoef` (the
        // plan's device copy of the same table) are not read any more.
        constexpr int NC = cfb_hc_ncoef(NS);
        const T *tab = HornerTable<T>::at(cfb_hc_off(NS));
        const T z = (T)(2 * (double)x1 + NS - 1.0);
        if (UNROLL) {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                T acc = tab[(NC - 1) * NS + i];
#pragma unroll
                for (int k = NC - 2; k >= 0; --k) acc = fma(z, acc, tab[k * NS + i]);
                dst[i] = acc;
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < NS; ++i) {
                T acc = tab[(NC - 1) * NS + i];
#pragma unroll
                for (int k = NC - 2; k >= 0; --k) acc = fma(z, acc, tab[k * NS + i]);
                dst[i] = acc;
            }
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
