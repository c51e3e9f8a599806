"""Low-level ctypes binding of libcufinufft.so (B200 build).

Plays the role of the reference's python/cufinufft/_cufinufft.py:36-153 -- same
exported names (`NufftOpts`, `_default_opts`, `_make_plan[f]`, `_set_pts[f]`,
`_exec_plan[f]`, `_destroy_plan[f]`) with the same argument order -- but
 * it needs no `imp` (absent from Python 3.12),
 * it loads the in-tree library  <package>/lib/libcufinufft.so  first (the driver
   records which .so files were loaded), then $CUFINUFFT_B200_LIB, then the
   dynamic loader's `libcufinufft.so` as the reference does (:39), and
 * it FAILS LOUDLY if none is found: there is no CPU or pure-Python fallback.
The extension symbols of include/cufinufft_b200.h are bound here as well.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_void_p

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = None


def _find_library():
    global LIB_PATH
    tried = []
    for cand in (os.path.join(_PKG_DIR, "lib", "libcufinufft.so"), os.environ.get("CUFINUFFT_B200_LIB"),
                 "libcufinufft.so"):
        if not cand:
            continue
        try:
            handle = ctypes.CDLL(cand)
            LIB_PATH = cand
            return handle
        except OSError as exc:
            tried.append("%s (%s)" % (cand, exc))
    raise RuntimeError(
        "Failed to find a suitable cufinufft library: build it with "
        "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C cufinufft_b200/csrc`.\n  tried: "
        + "\n         ".join(tried))


lib = _find_library()


class NufftOpts(Structure):
    """Field-for-field mirror of `cufinufft_opts` (include/cufinufft_opts.h), 64 bytes."""
    _fields_ = [("upsampfac", c_double)] + [(name, c_int) for name in (
        "gpu_method", "gpu_sort", "gpu_binsizex", "gpu_binsizey", "gpu_binsizez",
        "gpu_obinsizex", "gpu_obinsizey", "gpu_obinsizez", "gpu_maxsubprobsize",
        "gpu_nstreams", "gpu_kerevalmeth", "gpu_spreadinterponly", "gpu_device_id")]


NufftOpts_p = POINTER(NufftOpts)
c_int_p = POINTER(c_int)


def _get_ctypes(dtype):
    """float/double ctypes scalar and pointer types for a numpy real dtype."""
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return c_double, POINTER(c_double)
    if dtype == np.float32:
        return c_float, POINTER(c_float)
    raise TypeError("Expected np.float32 or np.float64.")


def _bind(name, argtypes, restype=c_int):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    return fn


_default_opts = _bind("cufinufft_default_opts", [c_int, c_int, NufftOpts_p])

# one set of entry points per precision: "" = double, "f" = single
_api = {}
for _sfx, _real in (("", c_double), ("f", c_float)):
    _rp = POINTER(_real)
    _api[_sfx] = dict(
        make_plan=_bind("cufinufft%s_makeplan" % _sfx,
                        [c_int, c_int, c_int_p, c_int, c_int, _real, c_int, POINTER(c_void_p), NufftOpts_p]),
        set_pts=_bind("cufinufft%s_setpts" % _sfx, [c_int, c_void_p, c_void_p, c_void_p, c_int, _rp, _rp, _rp, c_void_p]),
        exec_plan=_bind("cufinufft%s_execute" % _sfx, [c_void_p, c_void_p, c_void_p]),
        destroy_plan=_bind("cufinufft%s_destroy" % _sfx, [c_void_p]),
        # extensions (include/cufinufft_b200.h)
        set_stream=_bind("cufinufft%s_set_stream" % _sfx, [c_void_p, c_void_p]),
        set_pts_host=_bind("cufinufft%s_setpts_host" % _sfx, [c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
        exec_host=_bind("cufinufft%s_execute_host" % _sfx, [c_void_p, c_void_p, c_void_p]),
        spread=_bind("cufinufft%s_spread" % _sfx, [c_void_p, c_void_p, c_int, c_void_p]),
        interp=_bind("cufinufft%s_interp" % _sfx, [c_void_p, c_void_p, c_int, c_void_p]),
        get_ints=_bind("cufinufft%s_get_ints" % _sfx, [c_void_p, c_int, c_void_p]),
        get_reals=_bind("cufinufft%s_get_reals" % _sfx, [c_void_p, c_int, c_void_p]),
        set_timing=_bind("cufinufft%s_set_timing" % _sfx, [c_void_p, c_int]),
        get_timing=_bind("cufinufft%s_get_timing" % _sfx, [c_void_p, c_void_p]),
        get_launch_counts=_bind("cufinufft%s_get_launch_counts" % _sfx, [c_void_p, c_void_p]),
        set_interp_engine=_bind("cufinufft%s_set_interp_engine" % _sfx, [c_void_p, c_int]),
        set_sort_levels=_bind("cufinufft%s_set_sort_levels" % _sfx, [c_void_p, c_int]),
        # z-slab decomposition of one 3-D transform
        slab_make_plan=_bind("cufinufft%s_slab_makeplan" % _sfx,
                             [c_int, c_int_p, c_int, _real, c_int, c_int, POINTER(c_void_p), NufftOpts_p]),
        slab_info=_bind("cufinufft%s_slab_info" % _sfx, [c_void_p, c_void_p]),
        slab_type2=_bind("cufinufft%s_slab_type2" % _sfx, [c_void_p, c_void_p, c_void_p]),
        slab_type1_spread=_bind("cufinufft%s_slab_type1_spread" % _sfx, [c_void_p, c_void_p]),
        slab_halo_pack=_bind("cufinufft%s_slab_halo_pack" % _sfx, [c_int, c_void_p, c_void_p]),
        slab_halo_add=_bind("cufinufft%s_slab_halo_add" % _sfx, [c_int, c_void_p, c_void_p]),
        slab_type1_finish=_bind("cufinufft%s_slab_type1_finish" % _sfx, [c_void_p, c_void_p]),
        # the same with the collectives inside the library (NCCL; csrc/mgpu.cu)
        slab_set_comm=_bind("cufinufft%s_slab_set_comm" % _sfx, [c_void_p, c_void_p]),
        slab_route_setpts=_bind("cufinufft%s_slab_route_setpts" % _sfx, [c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
        slab_route_info=_bind("cufinufft%s_slab_route_info" % _sfx, [c_void_p, c_void_p]),
        slab_route_forward=_bind("cufinufft%s_slab_route_forward" % _sfx, [c_void_p, c_void_p, c_void_p]),
        slab_route_backward=_bind("cufinufft%s_slab_route_backward" % _sfx, [c_void_p, c_void_p, c_void_p]),
        slab_execute=_bind("cufinufft%s_slab_execute" % _sfx, [c_void_p, c_void_p, c_void_p]),
    )

# reference-compatible module-level names
_make_plan, _make_planf = _api[""]["make_plan"], _api["f"]["make_plan"]
_set_pts, _set_ptsf = _api[""]["set_pts"], _api["f"]["set_pts"]
_exec_plan, _exec_planf = _api[""]["exec_plan"], _api["f"]["exec_plan"]
_destroy_plan, _destroy_planf = _api[""]["destroy_plan"], _api["f"]["destroy_plan"]

version = _bind("cufinufft_b200_version", [], c_char_p)
host_params = _bind("cufinufft_b200_host_params", [c_int, c_int, c_int_p, c_double, c_int, NufftOpts_p, c_int_p,
                                                   POINTER(c_double)])
host_workplan = _bind("cufinufft_b200_host_workplan", [c_int, c_int, c_int_p, c_double, c_int, NufftOpts_p, ctypes.c_longlong, c_int_p])
phihat_quadrature = _bind("cufinufft_b200_phihat_quadrature", [c_int, c_int, c_double, c_double, c_double, c_int,
                                                               c_void_p, c_void_p])

microbench = _bind("cufinufft_b200_microbench", [c_int, c_int, POINTER(c_double)])
mgpu_unique_id = _bind("cufinufft_mgpu_unique_id", [c_void_p])
mgpu_comm_create = _bind("cufinufft_mgpu_comm_create", [c_int, c_int, c_void_p, c_int, POINTER(c_void_p)])
mgpu_comm_destroy = _bind("cufinufft_mgpu_comm_destroy", [c_void_p])

C_ABI_SYMBOLS = [base % s for s in ("", "f") for base in (
    "cufinufft%s_default_opts", "cufinufft%s_makeplan", "cufinufft%s_setpts", "cufinufft%s_execute",
    "cufinufft%s_destroy")]
EXTENSION_SYMBOLS = ["cufinufft_b200_version", "cufinufft_b200_host_params", "cufinufft_b200_host_workplan", "cufinufft_b200_phihat_quadrature", "cufinufft_b200_microbench",
                     "cufinufft_mgpu_unique_id", "cufinufft_mgpu_comm_create", "cufinufft_mgpu_comm_destroy"] + [base % s for s in ("", "f") for base in (
    "cufinufft%s_set_stream", "cufinufft%s_setpts_host", "cufinufft%s_execute_host", "cufinufft%s_spread",
    "cufinufft%s_interp", "cufinufft%s_get_ints", "cufinufft%s_get_reals", "cufinufft%s_set_timing",
    "cufinufft%s_get_timing", "cufinufft%s_get_launch_counts", "cufinufft%s_set_interp_engine", "cufinufft%s_set_sort_levels",
    "cufinufft%s_slab_makeplan", "cufinufft%s_slab_info", "cufinufft%s_slab_type2", "cufinufft%s_slab_type1_spread",
    "cufinufft%s_slab_halo_pack", "cufinufft%s_slab_halo_add", "cufinufft%s_slab_type1_finish",
    "cufinufft%s_slab_set_comm", "cufinufft%s_slab_route_setpts", "cufinufft%s_slab_route_info", "cufinufft%s_slab_route_forward",
    "cufinufft%s_slab_route_backward", "cufinufft%s_slab_execute")]
