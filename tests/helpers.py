"""Shared helpers for the parity tests (seeded synthetic inputs, GPU round trips)."""
import numpy as np


def cdtype(dtype):
    return np.complex64 if np.dtype(dtype) == np.float32 else np.complex128


def make_points(M, dim, dtype, seed=0, dist="uniform"):
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        pts = [rng.uniform(-np.pi, np.pi, M) for _ in range(dim)]
    elif dist == "cluster":           # Gaussian blobs folded into [-pi, pi)
        cen = rng.uniform(-np.pi, np.pi, (4, dim))
        which = rng.integers(0, 4, M)
        pts = [np.mod(cen[which, d] + 0.05 * 2 * np.pi * rng.standard_normal(M) + np.pi, 2 * np.pi) - np.pi
               for d in range(dim)]
    elif dist == "wide":              # the full valid range [-3pi, 3pi)
        pts = [rng.uniform(-3 * np.pi, 3 * np.pi, M) for _ in range(dim)]
    elif dist == "onebin":            # the reference spread tests' "all in one bin" case
        pts = [np.pi * rng.uniform(0, 1, M) / 64 for _ in range(dim)]
    else:
        raise ValueError(dist)
    return [np.ascontiguousarray(p.astype(dtype)) for p in pts]


def make_strengths(M, dtype, seed=1, ntransf=1):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (ntransf, M)) + 1j * rng.uniform(-1, 1, (ntransf, M))
    return np.ascontiguousarray(c.astype(cdtype(dtype)))


def make_modes_data(modes, dtype, seed=2, ntransf=1):
    rng = np.random.default_rng(seed)
    shape = (ntransf,) + tuple(modes)[::-1]
    fk = rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)
    return np.ascontiguousarray(fk.astype(cdtype(dtype)))


def rel_l2(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def gpu_nufft(nufft_type, modes, pts, data, tol, dtype, ntransf=1, maxbatch=1, iflag=None, return_plan=False,
              sort_levels=0, **opts):
    """Run our library through the Python class (C ABI underneath). modes x-fastest (ms,mt,mu);
    pts = [x,y,z]; data [ntransf][...]."""
    from cufinufft_b200 import cufinufft, gpuarray
    shape = tuple(modes)[::-1]
    plan = cufinufft(nufft_type, shape, n_trans=ntransf, eps=tol, isign=iflag, dtype=dtype, maxbatch=maxbatch, **opts)
    dev = [gpuarray.to_gpu(p) for p in pts]
    if sort_levels:
        plan.set_sort_levels(sort_levels)
    plan.set_pts(*dev[::-1])
    M = pts[0].size
    cd = cdtype(dtype)
    if nufft_type == 1:
        cg = gpuarray.to_gpu(np.ascontiguousarray(data, cd))
        fkg = gpuarray.zeros((ntransf,) + shape, cd)
        plan.execute(cg, fkg)
        out = fkg.get()
    else:
        fkg = gpuarray.to_gpu(np.ascontiguousarray(data, cd))
        cg = gpuarray.zeros((ntransf, max(M, 1)), cd)
        plan.execute(cg, fkg)
        out = cg.get()[:, :M]
    if return_plan:
        return out, plan
    return out


def drop_exact_stencil_points(pts, nf, ns):
    """Remove points whose rescaled coordinate makes `x_r - ns/2` an exact integer in any
    dimension.  For those the reference loops over ns+1 stencil points and reads ker[ns],
    an UNINITIALISED local (src/2d/spreadinterp2d.cu:35-38 with src/cuspreadinterp.h:37;
    SURVEY.md A.1), so its own output is garbage there (measured: such points alone move its
    result by up to 2e-3 rel-l2 and it is then further from the direct sum than ours).
    Parity against the reference library is therefore asserted on inputs without them.
    Rescale as contrib/spreadinterp.h:36-38 (double arithmetic, narrowed to the real type)."""
    dtype = pts[0].dtype.type
    keep = np.ones(pts[0].size, bool)
    pi = dtype(np.pi)
    for d, x in enumerate(pts):
        shift = np.where(x < -pi, 1.5, np.where(x >= pi, -0.5, 0.5))
        xr = ((x.astype(np.float64) * 0.159154943091895336 + shift) * nf[d]).astype(dtype).astype(np.float64)
        t = xr - ns / 2.0
        keep &= np.ceil(t) != t
    return [np.ascontiguousarray(p[keep]) for p in pts]
