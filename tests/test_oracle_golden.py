"""The oracle (oracle/xline_oracle.py) against (1) the committed outputs of the reference's
own element code (tests/golden, made by oracle/make_golden.py) and (2) the known-answer
values the reference's own tests hold (reference tests/test_beamfields.py, test_track.py,
test_losses.py, test_qgauss.py)."""
import numpy as np
import pytest

from oracle import xline_oracle as xo
from tests import helpers as H

CASES = sorted(H.manifest().keys())

# elements whose arithmetic is only + - * / sqrt and library sin/cos/exp/wofz evaluated by
# the same NumPy/SciPy build: the vectorised restatement must reproduce the reference's
# scalar/np.vectorize evaluation to the last bit.
BIT_EXACT_TOL = 0.0
# np.vectorize'd scalar math (math-library scalar vs SIMD loops may differ in the last ulp)
ULP_TOL = 4e-16


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_outputs(case):
    m, specs, cols, ref = H.load_case(case)
    got = H.run_oracle(specs, cols, m["p0c"], m["mass0"], num_turns=m.get("num_turns", 1))
    assert np.array_equal(got["state"], ref["state"])
    assert np.array_equal(got["at_element"], ref["at_element"])
    assert np.array_equal(got["at_turn"], ref["at_turn"])
    for k in H.COORDS + ("rpp", "rvv", "s"):
        err = H.rel_err(got[k], ref[k])
        tol = ULP_TOL if any(t in case for t in ("bb", "sc_", "rfmult", "cavity", "line")) else BIT_EXACT_TOL
        assert err <= tol, (case, k, err)


def _one(name, fields, p0c=1e9, mass0=938.272046e6, **cols):
    n = max([len(np.atleast_1d(v)) for v in cols.values()] + [1])
    p = xo.OracleParticles(n, p0c=p0c, mass0=mass0, **cols)
    xo.line_track([(name, fields)], p, 1)
    return p


def test_kat_spacecharge_reference_tests_beamfields():
    """reference tests/test_beamfields.py:9-83 (np.isclose default rtol=1e-5, atol=1e-15)."""
    x_co, y_co, sx, sy = 0.1, -0.5, 0.5, 0.1
    f1 = dict(number_of_particles=1e11, bunchlength_rms=0.22, sigma_x=sx, sigma_y=sy,
              length=2.0, x_co=x_co, y_co=y_co)
    f2 = dict(number_of_particles=1e11, circumference=0.22 * np.sqrt(2 * np.pi), sigma_x=sx,
              sigma_y=sy, length=2.0, x_co=x_co, y_co=y_co)
    p1 = _one("SCQGaussProfile", f1, x=x_co + 0.2, y=y_co - 0.5)
    p2 = _one("SCCoasting", f2, x=x_co + 0.2, y=y_co - 0.5)
    assert np.isclose(p1.px[0], 1.8329795395186613e-07, atol=1e-15)
    assert np.isclose(p1.py[0], -8.540420459001383e-07, atol=1e-15)
    assert abs(p1.px[0] - p2.px[0]) < 1e-15 and abs(p1.py[0] - p2.py[0]) < 1e-15
    # swapped axes (:49-59)
    f1s = dict(f1, sigma_x=sy, sigma_y=sx)
    p1 = _one("SCQGaussProfile", f1s, x=x_co - 0.5, y=y_co + 0.2)
    assert np.isclose(p1.px[0], -8.540420459001383e-07, atol=1e-15)
    assert np.isclose(p1.py[0], 1.8329795395186613e-07, atol=1e-15)
    # on the closed orbit (:61-70)
    p1 = _one("SCQGaussProfile", f1s, x=x_co, y=y_co)
    assert abs(p1.px[0]) < 1e-15 and abs(p1.py[0]) < 1e-15
    # round beam (:72-83)
    f1r = dict(f1s, sigma_y=sy)
    p1 = _one("SCQGaussProfile", f1r, x=x_co + 0.5, y=y_co + 0.1)
    assert np.isclose(p1.px[0], 1.2895332740238447e-06, atol=1e-15)
    assert np.isclose(p1.py[0], 2.579066548047689e-07, atol=1e-15)


def test_ellip_equal_sigmas_raises():
    """reference tests/test_beamfields.py:86-98."""
    with pytest.raises(ZeroDivisionError):
        xo.field_gauss_ellip(1.0, 1.0, np.array([0.5]), np.array([0.1]))


def test_rfmultipole_equals_multipole_at_zero_frequency():
    """reference tests/test_track.py:33-45 (abs_tol=1e-15)."""
    knl, ksl = [0.5, 2, 0.2], [0.5, 3, 0.1]
    p1 = _one("RFMultipole", dict(knl=knl, ksl=ksl), x=1.0, y=1.0)
    p2 = _one("Multipole", dict(knl=knl, ksl=ksl), x=1.0, y=1.0)
    for k in H.COORDS:
        assert abs(getattr(p1, k)[0] - getattr(p2, k)[0]) <= 1e-15


@pytest.mark.parametrize(
    "name,fields,mask",
    [
        ("LimitRect", dict(min_x=-0.1, max_x=0.3, min_y=-0.5, max_y=0.1),
         lambda x, y: (x >= -0.1) & (x <= 0.3) & (y >= -0.5) & (y <= 0.1)),
        ("LimitEllipse", dict(a=0.1, b=0.2), lambda x, y: x ** 2 / 0.1 ** 2 + y ** 2 / 0.2 ** 2 <= 1.0),
        ("LimitRectEllipse", dict(max_x=0.1, max_y=0.05, a=0.1, b=0.2),
         lambda x, y: (x ** 2 / 0.1 ** 2 + y ** 2 / 0.2 ** 2 <= 1.0) & (x >= -0.1) & (x <= 0.1)
         & (y >= -0.05) & (y <= 0.05)),
    ],
)
def test_apertures_reference_tests_track(name, fields, mask):
    """reference tests/test_track.py:48-129: survivors == NumPy mask, then everything lost."""
    arr = np.arange(0, 1, 0.001)
    p = _one(name, fields, x=arr, y=arr)
    assert len(p) == int(mask(arr, arr).sum())
    p.x = p.x + 0.3 + 1e-6
    xo.line_track([(name, fields)], p, 1)
    assert len(p) == 0
    full = xo.gather_full(p, len(arr))
    assert (full["state"] == 0).all()
    p = _one(name, fields, x=1.0, y=1.0)
    assert len(p) == 0


def test_particle_loss_compaction():
    """reference tests/test_losses.py:5-17."""
    p = xo.OracleParticles(10, p0c=1e9, x=np.arange(10, dtype=np.float64))
    p.state = np.int64(1) * (np.mod(p.x.astype(np.int64), 2) == 0)
    p.remove_lost_particles()
    p.state = np.int64(1) * (p.x > 5)
    p.remove_lost_particles()
    assert np.all(p.x == np.array([6.0, 8.0]))


def test_particle_reference_energy():
    """reference tests/test_particles.py:9-25 (p0c^2 + mass0^2 == energy0^2)."""
    for p0c in (1e9, 0.1 * 938.27208816e6):
        p = xo.OracleParticles(1, p0c=p0c)
        err = abs(p.p0c ** 2 + p.mass0 ** 2 - p.energy0 ** 2) / p.mass0 ** 2
        assert err < 1e-15


def test_qgauss_q1_is_gaussian():
    """reference tests/test_qgauss.py:6-29."""
    assert np.allclose(xo.qgauss_cq(1.0), np.sqrt(np.pi), 1e-16, 1e-16)
    for sigma in (1.0, 2.37):
        x = np.linspace(-4 * sigma, 4 * sigma, 101)
        ref = np.exp(-(x / sigma) ** 2 / 2.0) / np.sqrt(2 * np.pi * sigma * sigma)
        got = xo.qgauss_eval(x, 1.0, 1 / (np.sqrt(2) * sigma), xo.qgauss_cq(1.0))
        assert np.allclose(ref, got, 1e-15, 1e-16)


def test_monitor_store_index():
    f = dict(num_stores=3, start=2, skip=2, is_rolling=False)
    assert [xo.monitor_store_index(f, t) for t in range(10)] == [-1, -1, 0, -1, 1, -1, 2, -1, -1, -1]
    f["is_rolling"] = True
    assert xo.monitor_store_index(f, 8) == 0


@pytest.mark.parametrize("p0c,mass0", [(450e9, 938.27208816e6), (0.571e9, 938.27208816e6), (6e9, 0.51099895e6)])
def test_restated_energy_bookkeeping_is_relativistic_kinematics(p0c, mass0):
    """`add_to_energy` and the `delta` setter belong to the unvendored `xpart` container and are
    restated from memory (the part of the oracle the reference's own tests do not pin).  What
    they must compute is fixed by kinematics, independently of any implementation: with
    pc = p0c (1 + delta) and E = sqrt(pc^2 + m^2), adding dE gives E' = E + dE,
    pc' = sqrt(E'^2 - m^2), delta' = pc'/p0c - 1, rpp' = p0c/pc', rvv' = (pc'/E')/beta0 and
    zeta' = zeta rvv'/rvv.  Checked in 80-bit arithmetic for the oracle and for the host-side
    `xline_b200.Particles` (protons at 450 GeV and at PS Booster energy, electrons at 6 GeV)."""
    import torch

    import xline_b200 as xl

    L = np.longdouble
    rng = np.random.default_rng(5)
    n = 4000
    delta = rng.normal(0, 3e-3, n)
    zeta = rng.normal(0, 0.1, n)
    d_e = rng.normal(0, 1e-4, n) * p0c  # up to a few 1e-4 of the momentum per kick
    m, p0 = L(mass0), L(p0c)
    e0 = np.sqrt(p0 * p0 + m * m)
    beta0 = p0 / e0
    pc = p0 * (1 + delta.astype(L))
    en = np.sqrt(pc * pc + m * m)
    rvv_before = (pc / en) / beta0
    en2 = en + d_e.astype(L)
    pc2 = np.sqrt(en2 * en2 - m * m)
    want = dict(delta=pc2 / p0 - 1, rpp=p0 / pc2, rvv=(pc2 / en2) / beta0)
    want["zeta"] = zeta.astype(L) * want["rvv"] / rvv_before

    o = xo.OracleParticles(n, p0c=p0c, mass0=mass0, zeta=zeta, delta=delta)
    assert np.max(np.abs(o.rvv - rvv_before.astype(np.float64))) <= 2e-15
    o.add_to_energy(d_e)
    p = xl.Particles(p0c=p0c, mass0=mass0, device="cpu", zeta=zeta, delta=delta)
    p.add_to_energy(torch.as_tensor(d_e))
    for k, w in want.items():
        w64 = w.astype(np.float64)
        scale = np.max(np.abs(w64)) if k in ("delta", "zeta") else 1.0
        # sqrt(1 + small) - 1 loses the bits of `small` below 1e-16: absolute, not relative, accuracy
        assert np.max(np.abs(getattr(o, k) - w64)) <= 4e-15 * max(scale, 1.0), k
        assert np.max(np.abs(getattr(p, k).numpy() - w64)) <= 4e-15 * max(scale, 1.0), k
