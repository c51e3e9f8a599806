/* include/cufinufft.h -- drop-in C ABI of the B200-native cuFINUFFT hot path.
 *
 * The ten symbols below are exactly the extern "C" entry points the reference
 * exports (include/cufinufft_eitherprec.h:311-324, macros at :101-245) and that
 * python/cufinufft/_cufinufft.py:113-153 binds with ctypes.  Plain pointers and
 * sizes only; all array arguments are DEVICE pointers on opts.gpu_device_id.
 *
 * Behavioural contract kept from the reference:
 *  - nmodes always has 3 readable ints; iflag>=0 means e^{+i k x};
 *    maxbatchsize==0 -> min(ntransf, 8); opts==NULL -> defaults on device 0.
 *  - setpts borrows x,y,z until the next setpts/destroy; N,s,t,u are ignored.
 *  - c is [ntransf][M]; fk is [ntransf][mu][mt][ms] (x fastest), modes ordered
 *    -m/2 .. (m-1)/2.  execute() returns without synchronising the stream.
 *  - return 0 on success, non-zero on failure (codes in cufinufft_b200.h).
 * Deliberate differences: CUDA errors are returned as codes (the reference
 * calls exit(), contrib/cuda_samples/helper_cuda.h:583-590); a failed makeplan
 * stores NULL in *plan; destroy(NULL) returns 1 without dereferencing.      */
#ifndef CUFINUFFT_H_B200
#define CUFINUFFT_H_B200

#include <cuComplex.h>
#include "cufinufft_opts.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cufinufft_plan_s  *cufinufft_plan;   /* opaque (reference: public struct, used as void*) */
typedef struct cufinufftf_plan_s *cufinufftf_plan;

/* replaces CUFINUFFT_DEFAULT_OPTS, src/cufinufft.cu:639-730 */
int cufinufft_default_opts(int type, int dim, cufinufft_opts *opts);
int cufinufftf_default_opts(int type, int dim, cufinufft_opts *opts);

/* replaces CUFINUFFT_MAKEPLAN, src/cufinufft.cu:78-273 */
int cufinufft_makeplan(int type, int dim, int *nmodes, int iflag, int ntransf, double tol,
                       int maxbatchsize, cufinufft_plan *plan, cufinufft_opts *opts);
int cufinufftf_makeplan(int type, int dim, int *nmodes, int iflag, int ntransf, float tol,
                        int maxbatchsize, cufinufftf_plan *plan, cufinufft_opts *opts);

/* replaces CUFINUFFT_SETPTS, src/cufinufft.cu:275-493 */
int cufinufft_setpts(int M, double *x, double *y, double *z, int N, double *s, double *t, double *u,
                     cufinufft_plan plan);
int cufinufftf_setpts(int M, float *x, float *y, float *z, int N, float *s, float *t, float *u,
                      cufinufftf_plan plan);

/* replaces CUFINUFFT_EXECUTE, src/cufinufft.cu:495-569 */
int cufinufft_execute(cuDoubleComplex *c, cuDoubleComplex *fk, cufinufft_plan plan);
int cufinufftf_execute(cuFloatComplex *c, cuFloatComplex *fk, cufinufftf_plan plan);

/* replaces CUFINUFFT_DESTROY, src/cufinufft.cu:571-637 */
int cufinufft_destroy(cufinufft_plan plan);
int cufinufftf_destroy(cufinufftf_plan plan);

#ifdef __cplusplus
}
#endif
#endif
