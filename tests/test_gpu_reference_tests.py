"""The reference's own tests (tests/test_track.py, tests/test_beamfields.py) replayed through the
CUDA path with this package's classes: same constructions, same assertions and known answers.
Differences forced by the container semantics: lost particles stay in the arrays
with state == 0 until ``remove_lost_particles()`` is called, and scalars live in 1-element
tensors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_track_all():
    """tests/test_track.py:6-30: every element class, default-constructed, tracks a default
    particle without raising -- and, all strengths being zero, without moving it."""
    import xline_b200 as xl

    element_list = [xl.Drift, xl.DriftExact, xl.Multipole, xl.Cavity, xl.SawtoothCavity, xl.XYShift, xl.SRotation,
                    xl.RFMultipole, xl.BeamMonitor, xl.DipoleEdge, xl.Line, xl.LimitRect, xl.LimitEllipse,
                    xl.LimitRectEllipse, xl.BeamBeam4D, xl.BeamBeam6D, xl.SCCoasting, xl.SCQGaussProfile]
    for strict in (False, True):
        for el in element_list:
            p = xl.Particles(p0c=1e9)
            e = el()
            (e if isinstance(e, xl.Line) else xl.Line([e])).track(p, strict=strict)
            assert int(p.state[0]) == 1, el.__name__
            for k in ("x", "px", "y", "py", "zeta", "delta"):
                assert float(getattr(p, k)[0]) == 0.0, (el.__name__, k)


def test_track_rfmultipole():
    """tests/test_track.py:33-45."""
    import xline_b200 as xl

    p1 = xl.Particles(p0c=1e9, x=1, y=1)
    p2 = p1.copy()
    el1 = xl.RFMultipole(knl=[0.5, 2, 0.2], ksl=[0.5, 3, 0.1])
    el2 = xl.Multipole(knl=el1.knl, ksl=el1.ksl)
    el1.track(p1)
    el2.track(p2)
    assert p1.compare(p2, abs_tol=1e-15)


def test_track_LimitEllipse_and_LimitRectEllipse():
    """tests/test_track.py:76-129."""
    import xline_b200 as xl

    limit_a, limit_b, max_x, max_y = 0.1, 0.2, 0.1, 0.05
    arr = np.arange(0, 1, 0.001)
    cases = (
        (xl.LimitEllipse(a=limit_a, b=limit_b), arr ** 2 / limit_a ** 2 + arr ** 2 / limit_b ** 2 <= 1.0),
        (xl.LimitRectEllipse(max_x=max_x, max_y=max_y, a=limit_a, b=limit_b),
         (arr ** 2 / limit_a ** 2 + arr ** 2 / limit_b ** 2 <= 1.0) & (arr >= -max_x) & (arr <= max_x)
         & (arr >= -max_y) & (arr <= max_y)),
    )
    for el, survive in cases:
        p1 = xl.Particles(p0c=1e9, x=1, y=1)
        el.track(p1)
        assert int(p1.state[0]) == 0
        p2 = xl.Particles(x=arr, y=arr)
        el.track(p2)
        p2.remove_lost_particles()
        assert len(p2.state) == int(survive.sum())
        p2.x += limit_a + 1e-6
        el.track(p2)
        p2.remove_lost_particles()
        assert len(p2.x) == 0


@pytest.mark.parametrize("strict", [False, True])
def test_track_spacecharge(strict):
    """tests/test_beamfields.py:9-83: absolute kicks of the bunched and the coasting space-charge
    element for sigma_x > sigma_y, sigma_y > sigma_x, on the closed orbit and for a round beam."""
    import xline_b200 as xl

    x_co, y_co, sigma_x, sigma_y = 0.1, -0.5, 0.5, 0.1
    el1 = xl.SCQGaussProfile(number_of_particles=1e11, bunchlength_rms=0.22, sigma_x=sigma_x, sigma_y=sigma_y,
                             length=2.0, x_co=x_co, y_co=y_co)
    el2 = xl.SCCoasting(number_of_particles=el1.number_of_particles,
                        circumference=el1.bunchlength_rms * np.sqrt(2 * np.pi), sigma_x=el1.sigma_x,
                        sigma_y=el1.sigma_y, length=el1.length, x_co=el1.x_co, y_co=el1.y_co)

    def both(x, y):
        p1 = xl.Particles(p0c=1e9, x=x, y=y)
        p2 = p1.copy()
        xl.Line([el1]).track(p1, strict=strict)
        xl.Line([el2]).track(p2, strict=strict)
        assert p1.compare(p2, abs_tol=1e-15)
        return float(p1.px[0]), float(p1.py[0])

    x_offset, y_offset = 0.2, -0.5
    px, py = both(x_co + x_offset, y_co + y_offset)          # sigma_x > sigma_y
    assert np.isclose(px, 1.8329795395186613e-07, atol=1e-15) and np.isclose(py, -8.540420459001383e-07, atol=1e-15)
    for el in (el1, el2):
        el.sigma_x, el.sigma_y = sigma_y, sigma_x
    px, py = both(x_co + y_offset, y_co + x_offset)          # sigma_y > sigma_x
    assert np.isclose(px, -8.540420459001383e-07, atol=1e-15) and np.isclose(py, 1.8329795395186613e-07, atol=1e-15)
    px, py = both(el1.x_co, el1.y_co)                        # on the closed orbit
    assert np.isclose(px, 0.0, atol=1e-15) and np.isclose(py, 0.0, atol=1e-15)
    for el in (el1, el2):
        el.sigma_y = el.sigma_x                              # round beam
    px, py = both(el1.x_co + 0.5, el1.y_co + 0.1)
    assert np.isclose(px, 1.2895332740238447e-06, atol=1e-15) and np.isclose(py, 2.579066548047689e-07, atol=1e-15)


def test_get_transv_field_gauss_ellip_equal_sigmas():
    """tests/test_beamfields.py:86-98: the elliptical branch with equal sigmas divides by zero;
    here the error surfaces when the lattice is packed (min_sigma_diff = 0 forces that branch)."""
    import xline_b200 as xl

    with pytest.raises(ZeroDivisionError):
        xl.SCCoasting(number_of_particles=1.0, sigma_x=1.0, sigma_y=1.0, min_sigma_diff=0.0).track(
            xl.Particles(p0c=1e9, x=0.5, y=0.1))
