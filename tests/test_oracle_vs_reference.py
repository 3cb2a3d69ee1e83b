"""Pins the oracle against the REAL reference code, executed live.  Only runs where
/root/reference exists (the build container); on the GPU box it skips and the committed
fixtures in tests/golden (generated from the same code path) take over."""
import numpy as np
import pytest

from oracle import ref_harness as rh
from tests import helpers as H

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree absent")


def _fuzz_specs(rng):
    knl = list(rng.normal(0, 1, 6) * np.array([1e-4, 1e-2, 1, 10, 100, 1e4]))
    ksl = list(rng.normal(0, 1, 4) * np.array([1e-4, 1e-2, 1, 10]))
    return [
        ("Drift", dict(length=float(rng.uniform(0, 30)))),
        ("Multipole", dict(knl=knl, ksl=ksl, hxl=1e-3, hyl=-2e-4, length=2.0)),
        ("DriftExact", dict(length=float(rng.uniform(0, 30)))),
        ("SRotation", dict(angle=float(rng.uniform(-180, 180)))),
        ("XYShift", dict(dx=float(rng.normal(0, 1e-4)), dy=float(rng.normal(0, 1e-4)))),
        ("Cavity", dict(voltage=float(rng.uniform(0, 1e7)), frequency=4e8, lag=float(rng.uniform(0, 360)))),
        ("DipoleEdge", dict(h=0.01, e1=float(rng.uniform(-0.1, 0.1)), hgap=0.02, fint=0.5)),
        ("LimitEllipse", dict(a=3e-3, b=2e-3)),
        ("BeamBeam4D", dict(charge=1e11, sigma_x=float(rng.uniform(2e-4, 2e-3)),
                            sigma_y=float(rng.uniform(2e-4, 2e-3)), beta_r=1.0)),
        ("SCQGaussProfile", dict(number_of_particles=1e11, bunchlength_rms=0.2, sigma_x=1e-3,
                                 sigma_y=0.7e-3, length=1.0)),
        ("LimitRect", dict(min_x=-2e-3, max_x=2e-3, min_y=-2e-3, max_y=2e-3)),
    ]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz_line_two_turns(seed):
    from oracle.make_golden import beam, run_reference

    rng = np.random.default_rng(seed)
    specs = _fuzz_specs(rng)
    cols = beam(rng, 200, scale=1.5)
    ref = run_reference(specs, cols, 7e12, 938.27208816e6, num_turns=2)
    got = H.run_oracle(specs, cols, 7e12, 938.27208816e6, num_turns=2)
    for k in ("state", "at_element", "at_turn"):
        assert np.array_equal(got[k], ref[k]), k
    assert (ref["state"] == 0).sum() > 0
    for k in H.COORDS:
        assert H.rel_err(got[k], ref[k]) <= 4e-16, k


def test_golden_fixtures_are_current():
    """tests/golden was generated from this reference tree: re-run one case live."""
    from oracle.make_golden import run_reference

    m, specs, cols, out = H.load_case("multipole_ord5_skew")
    chi = cols.pop("chi")
    cols.pop("charge_ratio")
    ref = run_reference(specs, cols, m["p0c"], m["mass0"], chi=chi, charge_ratio=chi)
    for k in H.COORDS:
        assert np.array_equal(ref[k], out[k])


@pytest.mark.parametrize("seed", range(12))
def test_random_thin_lens_lines_bitwise(seed):
    """The random lines of the packer fuzz (tests/test_packed_format.py) through the REAL
    reference element code: the oracle must agree bit for bit on everything but the sin() of
    the cavities evaluated through the reference's scalar path (<= 4e-16)."""
    from oracle.make_golden import run_reference
    from tests.test_packed_format import _random_line

    rng = np.random.default_rng(1000 + seed)
    line = _random_line(rng)
    n = 150
    cols = dict(x=rng.normal(0, 8e-4, n), px=rng.normal(0, 1e-4, n), y=rng.normal(0, 8e-4, n),
                py=rng.normal(0, 1e-4, n), zeta=rng.normal(0, 0.05, n), delta=rng.normal(0, 3e-4, n))
    specs = line.to_specs()
    with np.errstate(all="ignore"):
        ref = run_reference(specs, cols, 26e9, 938.27208816e6, num_turns=3)
        got = H.run_oracle(specs, cols, 26e9, 938.27208816e6, num_turns=3)
    for k in ("state", "at_element", "at_turn"):
        assert np.array_equal(got[k], ref[k]), k
    has_cavity = any(name == "Cavity" for name, _ in specs)
    for k in H.COORDS:
        if has_cavity:
            assert H.rel_err(got[k], ref[k]) <= 1e-13, k  # a last-ulp sin() difference, carried through 3 turns
        else:
            assert np.array_equal(got[k], ref[k], equal_nan=True), k
