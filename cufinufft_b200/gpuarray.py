"""Minimal device array for tests / benchmarks: the subset of pycuda.gpuarray the
reference binding and its tests rely on (.ptr .dtype .size .shape, row views, .get(),
to_gpu()), implemented with ctypes on libcudart -- no pycuda, no torch.
"""
import ctypes
from ctypes import c_int, c_size_t, c_void_p

import numpy as np

_rt = None


def runtime():
    global _rt
    if _rt is None:
        last = None
        for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                _rt = ctypes.CDLL(name)
                break
            except OSError as exc:
                last = exc
        if _rt is None:
            raise RuntimeError("libcudart not found: %s" % last)
        _rt.cudaMalloc.argtypes = [ctypes.POINTER(c_void_p), c_size_t]
        _rt.cudaFree.argtypes = [c_void_p]
        _rt.cudaMemcpy.argtypes = [c_void_p, c_void_p, c_size_t, c_int]
        _rt.cudaMemset.argtypes = [c_void_p, c_int, c_size_t]
        _rt.cudaGetErrorString.restype = ctypes.c_char_p
        _rt.cudaGetErrorString.argtypes = [c_int]
    return _rt


def _check(code):
    if code != 0:
        raise RuntimeError("CUDA error %d: %s" % (code, runtime().cudaGetErrorString(code).decode()))


def device_count():
    n = c_int(0)
    code = runtime().cudaGetDeviceCount(ctypes.byref(n))
    return n.value if code == 0 else 0


def set_device(dev):
    _check(runtime().cudaSetDevice(c_int(int(dev))))


def synchronize():
    _check(runtime().cudaDeviceSynchronize())


class _Allocation:
    def __init__(self, nbytes):
        p = c_void_p()
        _check(runtime().cudaMalloc(ctypes.byref(p), max(int(nbytes), 1)))
        self.ptr = p.value

    def __del__(self):
        try:
            runtime().cudaFree(c_void_p(self.ptr))
        except Exception:
            pass


class GPUArray:
    def __init__(self, shape, dtype, _base=None, _offset=0):
        self.shape = (shape,) if np.isscalar(shape) else tuple(shape)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        self._alloc = _base if _base is not None else _Allocation(self.nbytes)
        self.ptr = self._alloc.ptr + _offset

    def __getitem__(self, i):
        """Leading-axis integer index -> contiguous view (what `k_gpu[0]` needs); a unit-stride
        slice of the leading axis -> contiguous view (what `k_gpu[1][:4]` needs)."""
        if isinstance(i, slice):
            start, stop, step = i.indices(self.shape[0])
            if step != 1:
                raise IndexError("only unit-stride slices are supported")
            n = max(stop - start, 0)
            row = (int(np.prod(self.shape[1:])) if len(self.shape) > 1 else 1) * self.dtype.itemsize
            return GPUArray((n,) + self.shape[1:], self.dtype, _base=self._alloc,
                            _offset=(self.ptr - self._alloc.ptr) + start * row)
        if not isinstance(i, (int, np.integer)) or len(self.shape) < 2:
            raise IndexError("only leading-axis integer views are supported")
        i = int(i) % self.shape[0]
        sub = self.shape[1:]
        step = int(np.prod(sub)) * self.dtype.itemsize
        return GPUArray(sub, self.dtype, _base=self._alloc, _offset=(self.ptr - self._alloc.ptr) + i * step)

    def set(self, host):
        host = np.ascontiguousarray(host, self.dtype)
        assert host.size == self.size
        _check(runtime().cudaMemcpy(c_void_p(self.ptr), host.ctypes.data_as(c_void_p), self.nbytes, 1))
        return self

    def get(self):
        out = np.empty(self.shape, self.dtype)
        _check(runtime().cudaMemcpy(out.ctypes.data_as(c_void_p), c_void_p(self.ptr), self.nbytes, 2))
        return out

    def fill_zero(self):
        _check(runtime().cudaMemset(c_void_p(self.ptr), 0, self.nbytes))
        return self


def to_gpu(host):
    host = np.ascontiguousarray(host)
    return GPUArray(host.shape, host.dtype).set(host)


def zeros(shape, dtype):
    return GPUArray(shape, dtype).fill_zero()


def empty(shape, dtype):
    return GPUArray(shape, dtype)
