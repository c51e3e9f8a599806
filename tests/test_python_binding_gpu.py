"""The reference's Python-level acceptance checks, re-stated against OUR binding through the
drop-in import path (`from cufinufft import cufinufft` with cufinufft_b200/compat on sys.path) and
our GPUArray shim in pycuda's place.  What is checked, and where the reference checks it:
  * 3-D type 1 / type 2, both precisions: relative error < 0.01 against the direct sum at flat index
    int(0.1789 * prod(shape)) / at one target, seed-0 Gaussian data
    (python/cufinufft/tests/test_basic.py:13-101, utils.py:26-89);
  * many transforms at once: error < 10 * eps (examples/example2d1many.py:47-65, example2d2many.py);
  * TypeError on wrong coordinate / data dtypes and on ragged coordinate arrays with the reference's
    message substrings, TypeError "Invalid option" on unknown keyword
    (python/cufinufft/tests/test_error_checks.py:12-101);
  * a plan per device with gpu_device_id (python/cufinufft/tests/test_multi.py:13-66)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api():
    compat = os.path.join(ROOT, "cufinufft_b200", "compat")
    sys.path.insert(0, compat)
    try:
        from cufinufft import cufinufft          # the name a user of the reference imports
        from cufinufft_b200 import gpuarray
        yield cufinufft, gpuarray
    finally:
        sys.path.remove(compat)
        sys.modules.pop("cufinufft", None)


def _points(M, dim, dtype):
    rng = np.random.RandomState(0)
    return rng.uniform(-np.pi, np.pi, (dim, M)).astype(dtype)


def _gauss(n, cdtype):
    rng = np.random.RandomState(0)
    v = rng.standard_normal(2 * n)
    return (v[0::2] + 1j * v[1::2]).astype(cdtype)


def _mode_grid(shape):
    axes = [np.arange(-(n // 2), (n + 1) // 2) for n in shape]
    return np.stack([g.ravel() for g in np.meshgrid(*axes, indexing="ij")])       # [dim][prod(shape)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_type1_one_mode_against_direct_sum(api, dtype):
    cufinufft, gpuarray = api
    cd = np.complex64 if dtype == np.float32 else np.complex128
    shape, M, tol = (16, 16, 16), 4096, 1e-3
    k = _points(M, 3, dtype)
    c = _gauss(M, cd)
    k_gpu, c_gpu = gpuarray.to_gpu(k), gpuarray.to_gpu(c)
    fk_gpu = gpuarray.GPUArray(shape, dtype=cd)
    plan = cufinufft(1, shape, eps=tol, dtype=dtype)
    plan.set_pts(k_gpu[0], k_gpu[1], k_gpu[2])
    plan.execute(c_gpu, fk_gpu)
    ind = int(0.1789 * np.prod(shape))
    want = np.sum(c * np.exp(1j * (_mode_grid(shape)[:, ind] @ k.astype(np.float64))))
    got = fk_gpu.get().ravel()[ind]
    assert abs(got - want) / abs(want) < 0.01


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_type2_one_target_against_direct_sum(api, dtype):
    cufinufft, gpuarray = api
    cd = np.complex64 if dtype == np.float32 else np.complex128
    shape, M, tol = (16, 16, 16), 4096, 1e-3
    k = _points(M, 3, dtype)
    fk = _gauss(int(np.prod(shape)), cd).reshape(shape)
    k_gpu, fk_gpu = gpuarray.to_gpu(k), gpuarray.to_gpu(fk)
    c_gpu = gpuarray.GPUArray(shape=(M,), dtype=cd)
    plan = cufinufft(2, shape, eps=tol, dtype=dtype)
    plan.set_pts(k_gpu[0], k_gpu[1], k_gpu[2])
    plan.execute(c_gpu, fk_gpu)
    ind = M // 2
    want = np.sum(fk.ravel() * np.exp(-1j * (k[:, ind].astype(np.float64) @ _mode_grid(shape))))
    got = c_gpu.get()[ind]
    assert abs(got - want) / abs(want) < 0.01


@pytest.mark.parametrize("nufft_type", [1, 2])
def test_many_transforms_2d(api, nufft_type):
    cufinufft, gpuarray = api
    dtype, cd = np.float32, np.complex64
    shape, M, n_trans, eps = (24, 32), 3000, 4, 1e-4
    k = _points(M, 2, dtype)
    rng = np.random.RandomState(1)
    k_gpu = gpuarray.to_gpu(k)
    plan = cufinufft(nufft_type, shape, n_trans, eps=eps, dtype=dtype)
    plan.set_pts(k_gpu[0], k_gpu[1])
    grid = _mode_grid(shape)
    if nufft_type == 1:
        c = (rng.standard_normal((n_trans, M)) + 1j * rng.standard_normal((n_trans, M))).astype(cd)
        fk_gpu = gpuarray.GPUArray((n_trans,) + shape, dtype=cd)
        plan.execute(gpuarray.to_gpu(c), fk_gpu)
        fk = fk_gpu.get()
        for t in range(n_trans):
            ind = (t * 131 + 7) % int(np.prod(shape))
            want = np.sum(c[t] * np.exp(1j * (grid[:, ind] @ k.astype(np.float64))))
            assert abs(fk[t].ravel()[ind] - want) / np.abs(c[t]).sum() < 10 * eps
    else:
        fk = (rng.standard_normal((n_trans,) + shape) + 1j * rng.standard_normal((n_trans,) + shape)).astype(cd)
        c_gpu = gpuarray.GPUArray((n_trans, M), dtype=cd)
        plan.execute(c_gpu, gpuarray.to_gpu(fk))
        c = c_gpu.get()
        for t in range(n_trans):
            j = (t * 977 + 3) % M
            want = np.sum(fk[t].ravel() * np.exp(-1j * (k[:, j].astype(np.float64) @ grid)))
            assert abs(c[t, j] - want) / np.abs(fk[t]).sum() < 10 * eps


def test_wrong_coordinate_dtype_is_a_type_error(api):
    cufinufft, gpuarray = api
    k = _points(4096, 3, np.float32)
    good, bad = gpuarray.to_gpu(k), gpuarray.to_gpu(k.astype(np.float64))
    plan = cufinufft(1, (16, 16, 16), eps=1e-3, dtype=np.float32)
    for args in ((bad[0], good[1], good[2]), (good[0], bad[1], good[2]), (good[0], good[1], bad[2]), (bad[0], bad[1], bad[2])):
        with pytest.raises(TypeError):
            plan.set_pts(*args)


def test_ragged_coordinates_are_a_type_error(api):
    cufinufft, gpuarray = api
    k_gpu = gpuarray.to_gpu(_points(8, 3, np.float32))
    plan = cufinufft(1, (16, 16, 16), eps=1e-3, dtype=np.float32)
    with pytest.raises(TypeError) as err:
        plan.set_pts(k_gpu[0], k_gpu[1][:4])
    assert "kx and ky must be equal" in err.value.args[0]
    with pytest.raises(TypeError) as err:
        plan.set_pts(k_gpu[0], k_gpu[1], k_gpu[2][:4])
    assert "kx and kz must be equal" in err.value.args[0]


def test_unknown_option_is_a_type_error(api):
    cufinufft, _ = api
    with pytest.raises(TypeError) as err:
        cufinufft(1, (8, 8), foo="bar")
    assert "Invalid option 'foo'" in err.value.args[0]


def test_wrong_data_dtype_is_a_type_error(api):
    cufinufft, gpuarray = api
    shape, M = (16, 16, 16), 4096
    k_gpu = gpuarray.to_gpu(_points(M, 3, np.float32))
    c = _gauss(M, np.complex64)
    plan = cufinufft(1, shape, eps=1e-3, dtype=np.float32)
    plan.set_pts(k_gpu[0], k_gpu[1], k_gpu[2])
    with pytest.raises(TypeError):
        plan.execute(gpuarray.to_gpu(c), gpuarray.GPUArray(shape, dtype=np.complex128))
    with pytest.raises(TypeError):
        plan.execute(gpuarray.to_gpu(np.ascontiguousarray(c.real)), gpuarray.GPUArray(shape, dtype=np.complex64))


@pytest.mark.parametrize("opts", [dict(gpu_method=1), dict(gpu_method=2), dict(gpu_method=1, gpu_sort=0),
                                  dict(gpu_kerevalmeth=1), dict(gpu_maxsubprobsize=64), dict(gpu_binsizex=16, gpu_binsizey=16)])
def test_options_by_keyword(api, opts):
    cufinufft, gpuarray = api
    dtype, cd, shape, M = np.float32, np.complex64, (32, 40), 5000
    k = _points(M, 2, dtype)
    c = _gauss(M, cd)
    k_gpu = gpuarray.to_gpu(k)
    plan = cufinufft(1, shape, eps=1e-4, dtype=dtype, **opts)
    plan.set_pts(k_gpu[0], k_gpu[1])
    fk_gpu = gpuarray.GPUArray(shape, dtype=cd)
    plan.execute(gpuarray.to_gpu(c), fk_gpu)
    ind = int(0.1789 * np.prod(shape))
    want = np.sum(c * np.exp(1j * (_mode_grid(shape)[:, ind] @ k.astype(np.float64))))
    assert abs(fk_gpu.get().ravel()[ind] - want) / abs(want) < 0.01


def test_a_plan_on_every_device(api):
    cufinufft, gpuarray = api
    n = gpuarray.device_count()
    shape, M, dtype, cd = (16, 16, 16), 4096, np.float32, np.complex64
    k, c = _points(M, 3, dtype), _gauss(M, cd)
    ind = int(0.1789 * np.prod(shape))
    want = np.sum(c * np.exp(1j * (_mode_grid(shape)[:, ind] @ k.astype(np.float64))))
    for dev in range(n):
        gpuarray.set_device(dev)
        try:
            k_gpu, c_gpu = gpuarray.to_gpu(k), gpuarray.to_gpu(c)
            fk_gpu = gpuarray.GPUArray(shape, dtype=cd)
            plan = cufinufft(1, shape, eps=1e-3, dtype=dtype, gpu_device_id=dev)
            plan.set_pts(k_gpu[0], k_gpu[1], k_gpu[2])
            plan.execute(c_gpu, fk_gpu)
            assert abs(fk_gpu.get().ravel()[ind] - want) / abs(want) < 0.01
            plan.destroy()
        finally:
            gpuarray.set_device(0)
