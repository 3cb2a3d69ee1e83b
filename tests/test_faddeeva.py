"""The in-kernel Faddeeva algorithm (Weideman N=40, csrc/beamfields.cuh::wofz_q1), restated
in NumPy from the generated coefficient table, against scipy.special.wofz -- the function
the reference calls (xline/mathlibs.py:11-13)."""
import os
import re

import numpy as np
from scipy.special import wofz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table():
    txt = open(os.path.join(ROOT, "xline_b200", "csrc", "faddeeva_coeffs.inc")).read()
    L = float(re.search(r"XLB_WEID_L (\S+)", txt).group(1))
    body = txt.split("XLB_WEID_COEFFS")[1]
    coeffs = [float(t) for t in re.findall(r"[-+]?\d\.\d+e[-+]\d+|[-+]?\d+\.\d+", body)]
    return L, np.array(coeffs)


def weideman(z):
    L, a = _table()
    inv = 1.0 / (L - 1j * z)
    Z = (L + 1j * z) * inv
    p = np.zeros_like(Z) + a[0]
    for c in a[1:]:
        p = p * Z + c
    return 2 * p * inv * inv + inv / np.sqrt(np.pi)


def test_table_matches_generator():
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen", os.path.join(ROOT, "scripts", "gen_faddeeva_coeffs.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    L, a = gen.coefficients()
    L2, a2 = _table()
    assert len(a2) == 40 and L == L2 and np.array_equal(a, a2)


def test_weideman_matches_wofz_in_first_quadrant():
    rng = np.random.default_rng(0)
    r = 10 ** rng.uniform(-8, 6, 200_000)
    th = rng.uniform(0, np.pi / 2, 200_000)
    z = np.concatenate([
        r * np.exp(1j * th),
        rng.uniform(0, 8, 100_000) + 1j * rng.uniform(0, 8, 100_000),
        rng.uniform(0, 30, 50_000) + 1j * 10 ** rng.uniform(-12, -1, 50_000),
        1j * 10 ** rng.uniform(-8, 3, 10_000), 10 ** rng.uniform(-8, 1.5, 10_000) + 0j,
        np.array([0j, 1e-300 + 1e-300j]),
    ])
    ref = wofz(z)
    err = np.abs(weideman(z) - ref) / np.abs(ref)
    assert err.max() < 5e-14, (err.max(), z[err.argmax()])
