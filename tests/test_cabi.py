"""CPU tests (no GPU, no compute calls): the drop-in boundary.  libcufinufft.so loads, exports
every symbol include/*.h declares, the opts struct has the reference's layout, defaults follow
src/cufinufft.cu:639-730, and the product fails LOUDLY (no CPU fallback) when there is no device."""
import ctypes
import os
import re
import subprocess
from ctypes import byref, c_int, c_void_p

import numpy as np
import pytest

from conftest import HAS_GPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


def _declared_symbols():
    names = []
    for h in sorted(os.listdir(INCLUDE)):
        text = open(os.path.join(INCLUDE, h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(cufinufftf?_\w+)\s*\(", text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    from cufinufft_b200 import _cufinufft as ll
    declared = _declared_symbols()
    assert len(declared) >= 10 + 20
    for name in declared:
        assert hasattr(ll.lib, name), name
    # and the python-side lists cover exactly the headers
    assert sorted(ll.C_ABI_SYMBOLS + ll.EXTENSION_SYMBOLS) == declared
    # unmangled, dynamic: what `nm -D` shows a maintainer
    out = subprocess.run(["nm", "-D", "--defined-only", ll.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    for name in ll.C_ABI_SYMBOLS:
        assert name in exported, name


def test_library_is_the_in_tree_build():
    from cufinufft_b200 import _cufinufft as ll
    assert os.path.realpath(ll.LIB_PATH) == os.path.realpath(os.path.join(ROOT, "cufinufft_b200", "lib", "libcufinufft.so"))
    assert b"cufinufft-b200" in ll.version()


def test_opts_struct_layout_matches_reference():
    from cufinufft_b200._cufinufft import NufftOpts
    # include/cufinufft_opts.h:4-26 / python/cufinufft/_cufinufft.py:81-97: double + 13 ints -> 64 bytes
    assert ctypes.sizeof(NufftOpts) == 64
    names = [n for n, _ in NufftOpts._fields_]
    assert names == ["upsampfac", "gpu_method", "gpu_sort", "gpu_binsizex", "gpu_binsizey", "gpu_binsizez",
                     "gpu_obinsizex", "gpu_obinsizey", "gpu_obinsizez", "gpu_maxsubprobsize", "gpu_nstreams",
                     "gpu_kerevalmeth", "gpu_spreadinterponly", "gpu_device_id"]
    assert NufftOpts.gpu_method.offset == 8 and NufftOpts.gpu_device_id.offset == 56


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("nufft_type", [1, 2])
def test_default_opts(nufft_type, dim):
    from cufinufft_b200 import _cufinufft as ll
    o = ll.NufftOpts()
    assert ll._default_opts(nufft_type, dim, o) == 0
    assert o.upsampfac == 2.0 and o.gpu_sort == 1 and o.gpu_maxsubprobsize == 1024
    assert o.gpu_kerevalmeth == 0 and o.gpu_spreadinterponly == 0 and o.gpu_device_id == 0
    assert (o.gpu_binsizex, o.gpu_binsizey, o.gpu_binsizez) == (-1, -1, -1)
    assert o.gpu_method == (2 if nufft_type == 1 else 1)        # src/cufinufft.cu:679-716
    f = ll.lib.cufinufftf_default_opts
    f.argtypes = [c_int, c_int, ll.NufftOpts_p]
    o2 = ll.NufftOpts()
    assert f(nufft_type, dim, o2) == 0 and bytes(o2) == bytes(o)


def test_null_and_bad_handles():
    from cufinufft_b200 import _cufinufft as ll
    assert ll._destroy_plan(None) == 1 and ll._destroy_planf(None) == 1     # reference returns 1 (after a bad read)
    assert ll._exec_plan(None, None, None) != 0
    assert ll._set_pts(0, None, None, None, 0, None, None, None, None) != 0
    m = (c_int * 3)(8, 8, 1)
    assert ll._make_plan(1, 2, m, 1, 1, 1e-6, 1, None, None) != 0            # NULL plan pointer


@pytest.mark.skipif(HAS_GPU, reason="only meaningful where no device exists")
def test_no_device_means_loud_failure_not_a_fallback():
    from cufinufft_b200 import _cufinufft as ll
    from cufinufft_b200 import cufinufft
    m = (c_int * 3)(16, 16, 1)
    h = c_void_p(1234)
    ier = ll._make_planf(1, 2, m, 1, 1, 1e-4, 1, byref(h), None)
    assert ier == 11 and h.value is None          # CFB_ERR_CUDA, and *plan = NULL
    with pytest.raises(RuntimeError):
        cufinufft(1, (16, 16), eps=1e-4, dtype=np.float32)


def test_python_class_argument_checks_need_no_device():
    from cufinufft_b200 import cufinufft
    with pytest.raises(TypeError):
        cufinufft(1, (16, 16), dtype=np.int32)
    with pytest.raises(TypeError) as err:
        cufinufft(1, (16, 16), dtype=np.float32, not_an_option=1)
    assert "Invalid option" in err.value.args[0]
    with pytest.raises(RuntimeError):
        cufinufft(3, (16, 16), dtype=np.float32)              # 'Configuration not yet implemented.'


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under cufinufft_b200/ may import, link or exec it."""
    pkg = os.path.join(ROOT, "cufinufft_b200")
    for base, _, files in os.walk(pkg):
        if os.sep + "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")) or f == "Makefile":
                text = open(os.path.join(base, f), errors="replace").read()
                assert "oracle" not in text.replace("the oracle", ""), os.path.join(base, f)
    from cufinufft_b200 import _cufinufft as ll
    out = subprocess.run(["ldd", ll.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out


def test_compat_alias_package():
    import importlib
    import sys
    compat = os.path.join(ROOT, "cufinufft_b200", "compat")
    sys.path.insert(0, compat)
    try:
        mod = importlib.import_module("cufinufft")
        from cufinufft_b200 import cufinufft as ours
        assert mod.cufinufft is ours
    finally:
        sys.path.remove(compat)
        sys.modules.pop("cufinufft", None)
