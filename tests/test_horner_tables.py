"""The Horner coefficient tables (gpu_kerevalmeth=1): the product's and the oracle's copies hold
the REFERENCE's generated constants (contrib/ker_horner_allw_loop.c:4-216), value for value.
CPU-only.  Where /root/reference exists (the build container) the committed tables are re-derived
from it; everywhere, the two copies must agree, approximate the ES kernel, and stay close to the
independent re-fit of csrc/gen_horner.py."""
import importlib.util
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT = os.path.join(ROOT, "cufinufft_b200", "csrc", "horner_coeffs.inc")
ORACLE = os.path.join(ROOT, "oracle", "horner_ref_table.inc")
REF = "/root/reference/contrib/ker_horner_allw_loop.c"


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def parse_inc(path):
    """{w: array[ncoef][w]} from one of the generated .inc files."""
    txt = open(path).read()
    ncoef = [int(v) for v in re.search(r"_horner_ncoef\[17\] = \{([^}]*)\}", txt).group(1).split(",")]
    out = {}
    for m in re.finditer(r"\{ /\* w=(\d+) \*/\n(.*?)\n  \},", txt, re.S):
        w = int(m.group(1))
        rows = [[float(v) for v in line.strip().rstrip(",").split(",")] for line in m.group(2).splitlines()]
        assert len(rows) == ncoef[w] and all(len(r) == 16 for r in rows)
        assert all(v == 0 for r in rows for v in r[w:])
        out[w] = np.array([r[:w] for r in rows])
    assert sorted(out) == list(range(2, 17))
    return out


def beta_of(w):
    return {2: 2.20, 3: 2.26, 4: 2.38}.get(w, 2.30) * w      # contrib/spreadinterp.cpp:58-66


def test_product_and_oracle_tables_are_the_same_numbers():
    a, b = parse_inc(PRODUCT), parse_inc(ORACLE)
    for w in range(2, 17):
        assert np.array_equal(a[w], b[w]), w


@pytest.mark.skipif(not os.path.exists(REF), reason="/root/reference not present (GPU box)")
def test_tables_equal_the_reference_table_bit_for_bit():
    imp = _load(os.path.join(ROOT, "tools", "import_horner_table.py"), "import_horner_table")
    ref = imp.parse(REF)
    for path in (PRODUCT, ORACLE):
        tab = parse_inc(path)
        for w in range(2, 17):
            want = np.array([[float(v) for v in row] for row in ref[w]])
            assert tab[w].shape == want.shape, (path, w)
            assert np.array_equal(tab[w], want), (path, w)


@pytest.mark.parametrize("w", range(2, 17))
def test_table_approximates_the_es_kernel(w):
    """Piecewise polynomial vs exp(beta*sqrt(1-(2x/w)^2)) (un-normalised, src/cuspreadinterp.h:6-16) on a
    fine grid of every interval, relative to the kernel peak e^beta: the fit error of the reference's
    table is below 10^(1-w) down to the fp64 floor."""
    tab = parse_inc(PRODUCT)[w]
    beta = beta_of(w)
    z = np.linspace(-1, 1, 201)
    worst = 0.0
    for i in range(w):
        x = -w / 2 + i + (z + 1) / 2
        exact = np.exp(beta * np.sqrt(np.maximum(0.0, 1 - (2 * x / w) ** 2)))
        poly = np.zeros_like(z)
        for k in range(tab.shape[0] - 1, -1, -1):
            poly = poly * z + tab[k, i]
        worst = max(worst, float(np.max(np.abs(poly - exact))) / np.exp(beta))
    assert worst <= max(10.0 ** (1 - w), 5e-13), worst


@pytest.mark.parametrize("w", range(2, 17))
def test_independent_refit_is_close(w):
    """csrc/gen_horner.py re-derives the table without the reference; documented deviation: interior
    intervals ~1e-14, edge intervals <= 4e-5 of the peak at w=2, 3e-6 (3), 3e-7 (>= 4), 2e-12 (>= 10)."""
    gen = _load(os.path.join(ROOT, "cufinufft_b200", "csrc", "gen_horner.py"), "gen_horner")
    tab = parse_inc(PRODUCT)[w]
    mine = gen.coeffs(w)
    assert mine.shape == tab.shape
    z = np.linspace(-1, 1, 101)
    dev = 0.0
    for i in range(w):
        pa = np.zeros_like(z)
        pb = np.zeros_like(z)
        for k in range(tab.shape[0] - 1, -1, -1):
            pa = pa * z + tab[k, i]
            pb = pb * z + mine[k, i]
        dev = max(dev, float(np.max(np.abs(pa - pb))))
    dev /= np.exp(beta_of(w))
    bound = {2: 6e-5, 3: 6e-6}.get(w, 6e-7 if w < 10 else 5e-12)
    assert dev <= bound, dev
