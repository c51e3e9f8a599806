/* include/cufinufft_opts.h -- options struct of the drop-in C ABI.
 *
 * Field order, types and total size (8 + 13*4 -> 64 bytes) are those of the
 * reference's include/cufinufft_opts.h:4-26 and of the ctypes mirror
 * python/cufinufft/_cufinufft.py:81-97; the struct cannot grow (Python
 * allocates it), so every extension goes through new symbols in
 * cufinufft_b200.h instead of new fields.                                   */
#ifndef CUFINUFFT_OPTS_H_B200
#define CUFINUFFT_OPTS_H_B200

typedef struct cufinufft_opts {
    double upsampfac;        /* upsampling ratio sigma; only 2.0 is implemented (as in the reference) */
    int gpu_method;          /* 1: NU-point driven (GM / GM-sort); 2: shared-memory subproblems (SM);
                                3: (2-D spread, reference "Paul") served by 2; 4: (3-D spread,
                                reference "block gather") obin-aligned fine grid, served by the SM engine */
    int gpu_sort;            /* method 1 only: 0 = input order (GM), 1 = bin-sorted (GM-sort) */
    int gpu_binsizex;        /* bin edge in fine-grid cells; <0 = default (1-D 1024; 2-D 32x32; 3-D 16x16x2, method 4: 4x4x4) */
    int gpu_binsizey;
    int gpu_binsizez;
    int gpu_obinsizex;       /* method 4 only: output-bin size (default 8); nf is rounded to a multiple of it */
    int gpu_obinsizey;
    int gpu_obinsizez;
    int gpu_maxsubprobsize;  /* max points per subproblem (default 1024) */
    int gpu_nstreams;        /* accepted and ignored (dead option in the reference too) */
    int gpu_kerevalmeth;     /* 0: exp(beta*sqrt(1-c x^2)); 1: Horner piecewise polynomial */
    int gpu_spreadinterponly;/* 0: full NUFFT; 1: spread (type 1) / interpolate (type 2) only, fk is the fine grid */
    int gpu_device_id;       /* CUDA device the plan lives on */
} cufinufft_opts;

#endif
