"""Host-only tests of the work plan `setpts` derives for the tile engines (csrc/spread.cu:
choose_internal_bins, plan_tile_geometry) through cufinufft_b200_host_workplan -- no device needed.
Invariants: sub-bins tile the reference's bins exactly (the reference-facing arrays are sums over
them), tiles fit the shared memory of a B200 SM, dense inputs get smaller tiles / larger work items,
sparse ones keep the reference's bins, type 2 plans are never split."""
from ctypes import c_int

import numpy as np
import pytest

KEYS = ("ibsx", "ibsy", "ibsz", "spbx", "spby", "spbz", "nibins", "nbins", "imaxsub", "ilist", "tile_cells", "tile_sy",
        "tile_sz", "sm_warps", "pad", "tile_cost")


def workplan(nufft_type, modes, tol, dtype, M, **opts):
    from cufinufft_b200 import _cufinufft as ll
    o = ll.NufftOpts()
    assert ll._default_opts(nufft_type, len(modes), o) == 0
    for k, v in opts.items():
        setattr(o, k, v)
    m = (c_int * 3)(*(tuple(modes) + (1,) * (3 - len(modes))))
    out = (c_int * 16)()
    ier = ll.host_workplan(nufft_type, len(modes), m, tol, int(np.dtype(dtype) == np.float32), o, int(M), out)
    assert ier == 0
    return dict(zip(KEYS, list(out)))


def _ref_bins(nufft_type, modes, dtype, opts):
    dim = len(modes)
    if dim == 1:
        d = [1024, 1, 1]
    elif dim == 2:
        d = [32, 32, 1]
    else:
        d = [16, 16, 2]
    for k, i in (("gpu_binsizex", 0), ("gpu_binsizey", 1), ("gpu_binsizez", 2)):
        if opts.get(k, -1) > 0 and i < dim:
            d[i] = opts[k]
    return d


CONFIGS = [
    # (type, modes, tol, dtype, M, opts)                                 BASELINE.json configs 1, 3, 4 and friends
    (1, (1000, 1000), 1e-3, np.float32, 10_000_000, {}),
    (1, (256, 256, 256), 1e-5, np.float32, 100_000_000, {}),
    (1, (512, 512), 1e-4, np.float32, 262_144, {}),
    (1, (512, 512, 512), 1e-9, np.float64, 1_000_000_000, {}),
    (1, (100, 80), 1e-6, np.float32, 300_000, dict(gpu_binsizex=24, gpu_binsizey=10)),
    (1, (64, 64, 64), 1e-12, np.float64, 5_000_000, {}),
    (1, (64, 48), 1e-4, np.float32, 100, {}),
    (1, (3000,), 1e-5, np.float32, 1_000_000, {}),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "t%d-%s-M%d-%s" % (c[0], "x".join(map(str, c[1])), c[4], np.dtype(c[3]).name))
def test_workplan_invariants(cfg):
    nufft_type, modes, tol, dtype, M, opts = cfg
    dim = len(modes)
    w = workplan(nufft_type, modes, tol, dtype, M, **opts)
    bs = _ref_bins(nufft_type, modes, dtype, opts)
    ibs, spb = [w["ibsx"], w["ibsy"], w["ibsz"]], [w["spbx"], w["spby"], w["spbz"]]
    for d in range(3):
        assert ibs[d] * spb[d] == bs[d], (d, ibs, spb, bs)                # sub-bins tile the reference bin exactly
        assert spb[d] & (spb[d] - 1) == 0                                  # by repeated halving
    assert w["nibins"] == w["nbins"] * spb[0] * spb[1] * spb[2]
    assert w["ilist"] == int(w["nibins"] != w["nbins"] or w["imaxsub"] != opts.get("gpu_maxsubprobsize", 1024))
    assert w["imaxsub"] >= opts.get("gpu_maxsubprobsize", 1024) and w["imaxsub"] <= max(4096, opts.get("gpu_maxsubprobsize", 1024))
    # the tile covers the internal bin + halo on every side, with padded strides
    ex, ey, ez = ibs[0] + 2 * w["pad"], (ibs[1] + 2 * w["pad"]) if dim > 1 else 1, (ibs[2] + 2 * w["pad"]) if dim > 2 else 1
    assert w["tile_sy"] >= ex
    if dim == 3:
        assert w["tile_sz"] >= w["tile_sy"] * ey
    assert w["tile_cells"] >= ex * ey * ez
    cell_bytes = 8 if dtype == np.float32 else 16
    ns = int(np.ceil(-np.log10(tol / 10.0)))
    plane = dtype == np.float64 and dim == 3 and nufft_type == 1 and ns >= 9 and opts.get("gpu_method", 2) == 2
    if plane and ns > 10:
        pass        # wider stencils: the engine applies only while its tile fits (ns = 13: 230 KB, the warp-private engine serves)
    elif plane:
        # the plane-owner engine (csrc/spread_plane.cuh): ONE tile per block = the whole reference bin, never split,
        # row stride = ns (mod 8) cells, a warp per tile plane; tile + the 256-point batch scratch fit the SM
        assert w["nibins"] == w["nbins"] and w["tile_sy"] % 8 == ns % 8 and w["sm_warps"] == min(ez, 16)
        assert w["tile_cells"] * cell_bytes + 256 * ((((4 * ns) // 2) | 1) * 2) * 8 + 4096 <= 227 * 1024
    elif w["sm_warps"] > 0:
        assert w["sm_warps"] * w["tile_cells"] * cell_bytes <= 227 * 1024
    if dim == 1:
        assert w["nibins"] == w["nbins"]                                   # 1-D is never split


def test_dense_inputs_get_more_resident_warps():
    # config 3: the reference's 22x22x8 tile allows 5 warps per SM; the split must at least double that
    w = workplan(1, (256, 256, 256), 1e-5, np.float32, 100_000_000)
    per_warp = w["tile_cells"] * 8
    assert w["nibins"] > w["nbins"] and per_warp <= 16 * 1024
    assert w["ibsx"] * w["ibsy"] * w["ibsz"] == 128                         # two halvings of the 16x16x2 bin
    assert w["imaxsub"] == 100_000_000 // (16 * 148 * 16)                 # larger work items (one tile flush each), >= 16 per resident warp
    # config 1: the 36x36 tile + the single-pass scratch already fit the 14 KB per-warp target; never
    # fewer than 64 points per (sub-)bin
    w1 = workplan(1, (1000, 1000), 1e-3, np.float32, 10_000_000)
    assert w1["tile_cells"] * 8 <= 14 * 1024 and 10_000_000 >= 64 * w1["nibins"]
    # config 4 (ns = 5: 38x38 tile): same
    w4 = workplan(1, (512, 512), 1e-4, np.float32, 262_144)
    assert w4["tile_cells"] * 8 <= 14 * 1024 and 262_144 >= 64 * w4["nibins"]


def test_sparse_inputs_and_type2_keep_the_reference_bins():
    w = workplan(1, (256, 256, 256), 1e-5, np.float32, 100_000)           # 0.4 points per bin
    assert w["nibins"] == w["nbins"] and w["ilist"] == 0
    w = workplan(2, (2048, 2048), 1e-9, np.float64, 40_000_000, gpu_method=1)
    assert w["nibins"] == w["nbins"] and w["ilist"] == 0 and w["imaxsub"] == 1024
    w = workplan(1, (1000, 1000), 1e-3, np.float32, 10_000_000, gpu_method=1)   # GM engine: no tiles, no split
    assert w["nibins"] == w["nbins"] and w["ilist"] == 0


def test_wide_fp64_stencils_are_split_until_four_warps_fit():
    # 3-D fp64 ns = 8 (the widest stencil the warp-private engine still serves in double precision): the
    # reference's bin needs a 92 KB tile (two warps per SM); the split goes on regardless of density while
    # fewer than four warps fit
    w = workplan(1, (512, 512, 512), 1e-7, np.float64, 1_000_000)
    assert w["nibins"] > w["nbins"]
    assert w["sm_warps"] >= 3 or min(w["ibsx"], w["ibsy"]) <= 4
    # ns = 10: the plane-owner engine takes over -- one block-shared 26 x 26 x 12 tile, no split
    w = workplan(1, (512, 512, 512), 1e-9, np.float64, 1_000_000)
    assert w["nibins"] == w["nbins"] and (w["ibsx"], w["ibsy"], w["ibsz"]) == (16, 16, 2)
    assert w["tile_sy"] == 26 and w["tile_sz"] == 26 * 26 and w["sm_warps"] == 12


def test_sm2_row_orders_are_permutations_and_conflict_free():
    """csrc/spread_sm2.cuh: sm2_row -- the compile-time row orders of the 3-D ns = 6 SM spread kernel.  Every order
    deals each of the 36 stencil rows to exactly one (pass, row slot); for the stride classes they were searched for
    (tools/search_sm2_rowmap.py) the 16 lanes of every half-warp of a run flush -- lane = (row slot, column pair), 8-byte
    cells, 16 bank pairs -- touch 16 different bank pairs, for both columns of the pair.  The strides are the ones the
    work plan picks for config 3's sub-bins (14, 197) and for the reference's own bins (22, 499)."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cufinufft_b200", "csrc", "spread_sm2.cuh")).read()
    body = src[src.index("constexpr signed char tab[SM2_RMC_COUNT][40] = {"):]
    body = body[:body.index("};")]
    tabs = [[int(v) for v in row.split(",")] for row in re.findall(r"\{([-\d,\s]+)\}", body)]
    assert len(tabs) == 3 and all(len(t) == 40 for t in tabs)
    for t in tabs:
        assert sorted(v for v in t if v >= 0) == list(range(36)) and t[36:] == [-1] * 4
    assert tabs[0][:36] == list(range(36))

    def conflicts(tab, sy, sz):
        worst = 0
        for it in range(4):
            for half in range(2):                                   # column xp, then column xp + 3
                for g0 in (0, 16):
                    seen = {}
                    for lane in range(g0, min(g0 + 16, 30)):
                        r, xp = divmod(lane, 3)
                        row = tab[it * 10 + r]
                        if row < 0:
                            continue
                        iz, iy = divmod(row, 6)
                        bank = (iz * sz + iy * sy + xp + 3 * half) % 16
                        seen[bank] = seen.get(bank, 0) + 1
                    worst = max(worst, max(seen.values(), default=1))
        return worst

    assert conflicts(tabs[1], 14, 197) == 1 and conflicts(tabs[2], 22, 499) == 1
    assert conflicts(tabs[0], 14, 196) == 2                       # the natural order on the natural strides: what round 1 measured
    # and the work plan picks exactly these strides
    w = workplan(1, (256, 256, 256), 1e-5, np.float32, 100_000_000)
    assert (w["tile_sy"], w["tile_sz"], w["tile_cost"]) == (14, 197, 16)
    w = workplan(1, (64, 64, 64), 1e-5, np.float32, 100_000)
    assert (w["tile_sy"], w["tile_sz"]) == (22, 499)
