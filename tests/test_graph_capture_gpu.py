"""`execute` is stream-ordered with no host synchronisation, no allocation and no host-side read-back, so a
caller can capture it into a CUDA graph on its own stream and replay it ("caller-supplied
stream, async execute, graph capture of the batch loop"; the reference blocks inside execute --
cudaDeviceSynchronize in type 2, src/2d/cufinufft2d.cu:139 -- and cannot be captured).  Captured here through
libcudart with ctypes: warm-up execute (buffers, cuFFT work area and kernel attributes exist afterwards),
capture one execute of a batch of transforms, replay twice with NEW data in the same buffers, compare with the
eager call."""
import ctypes
from ctypes import byref, c_int, c_ulonglong, c_void_p

import numpy as np
import pytest

from helpers import cdtype, make_modes_data, make_points, make_strengths, rel_l2

pytestmark = pytest.mark.gpu

CASES = [
    (1, (96, 80), 60_000, 1e-5, np.float32, 5, 2),       # three batches of <= 2 transforms inside the captured call
    (2, (96, 80), 60_000, 1e-9, np.float64, 3, 0),
    (1, (24, 20, 16), 40_000, 1e-5, np.float32, 1, 0),
    (2, (24, 20, 16), 40_000, 1e-9, np.float64, 2, 0),     # wide fp64 stencil: merged-tile interpolation
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "t%d-%s-%s-n%d" % (c[0], "x".join(map(str, c[1])), np.dtype(c[4]).name, c[5]))
def test_execute_can_be_captured_into_a_cuda_graph(case):
    from cufinufft_b200 import cufinufft, gpuarray
    rt = gpuarray.runtime()
    nufft_type, modes, M, tol, dtype, ntransf, maxbatch = case
    dim, shape, cd = len(modes), tuple(modes)[::-1], cdtype(dtype)
    pts = make_points(M, dim, dtype, seed=3)
    stream = c_void_p()
    assert rt.cudaStreamCreateWithFlags(byref(stream), 1) == 0          # cudaStreamNonBlocking
    plan = cufinufft(nufft_type, shape, n_trans=ntransf, eps=tol, dtype=dtype, maxbatch=maxbatch)
    plan.set_stream(stream.value)
    dev = [gpuarray.to_gpu(p) for p in pts]
    plan.set_pts(*dev[::-1])
    cg = gpuarray.zeros((ntransf, M), cd)
    fg = gpuarray.zeros((ntransf,) + shape, cd)

    def load(seed):
        if nufft_type == 1:
            cg.set(make_strengths(M, dtype, ntransf=ntransf, seed=seed))
        else:
            fg.set(make_modes_data(modes, dtype, ntransf=ntransf, seed=seed))

    def result():
        assert rt.cudaStreamSynchronize(stream) == 0
        return (fg if nufft_type == 1 else cg).get()

    load(1)
    plan.execute(cg, fg)                                               # warm-up, eager
    eager1 = result()
    load(2)
    plan.execute(cg, fg)
    eager2 = result()

    graph, gexec = c_void_p(), c_void_p()
    assert rt.cudaStreamBeginCapture(stream, c_int(0)) == 0             # cudaStreamCaptureModeGlobal: any illegal call fails the capture
    plan.execute(cg, fg)
    assert rt.cudaStreamEndCapture(stream, byref(graph)) == 0 and graph.value
    rt.cudaGraphInstantiate.argtypes = [ctypes.POINTER(c_void_p), c_void_p, c_ulonglong]
    assert rt.cudaGraphInstantiate(byref(gexec), graph, 0) == 0
    nnodes = ctypes.c_size_t(0)
    rt.cudaGraphGetNodes.argtypes = [c_void_p, c_void_p, ctypes.POINTER(ctypes.c_size_t)]
    assert rt.cudaGraphGetNodes(graph, None, byref(nnodes)) == 0 and nnodes.value >= 3

    tol_eq = 0 if nufft_type == 2 else 20 * np.finfo(dtype).eps         # type 1: atomic accumulation order
    for seed, want in ((1, eager1), (2, eager2)):
        load(seed)
        (cg if nufft_type == 2 else fg).fill_zero()
        assert rt.cudaGraphLaunch(gexec, stream) == 0
        got = result()
        if tol_eq == 0:
            assert np.array_equal(got, want)
        else:
            assert rel_l2(got, want) <= tol_eq
    rt.cudaGraphExecDestroy(gexec)
    rt.cudaGraphDestroy(graph)
    plan.destroy()
    rt.cudaStreamDestroy(stream)
