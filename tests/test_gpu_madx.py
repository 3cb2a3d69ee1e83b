"""GPU tests for the MAD-X import path: the reference's own
tests/test_madx_import.py error tests replayed through the CUDA tracker, the PS Booster
configuration C5 against the oracle, and the tracker as the check on xline_b200.optics."""
import math

import numpy as np
import pytest
import torch

from tests import helpers as H
from tests.test_gpu_parity import make_particles

pytestmark = pytest.mark.gpu


def test_neutral_errors():
    """tests/test_madx_import.py:173-237: misaligned (and re-aligned) collimators are
    transparent to a particle without transverse momentum."""
    import xline_b200 as xl
    from xline_b200.madx_input import MadxFile

    mad = MadxFile(text='''
        T1: Collimator, L=1.0, apertype=CIRCLE, aperture={0.5};
        T2: Collimator, L=1.0, apertype=CIRCLE, aperture={0.5};
        T3: Collimator, L=1.0, apertype=CIRCLE, aperture={0.5};
        KQ1 = 0.02;
        KQ2 = -0.02;
        testseq: SEQUENCE, l = 20.0;
            T1, at =  5;
            T2, at = 12;
            T3, at = 18;
        ENDSEQUENCE;
        BEAM, PARTICLE=PROTON, ENERGY=7000.0, EXN=2.2e-6, EYN=2.2e-6;
        USE, SEQUENCE=testseq;
        Select, flag=makethin, pattern="T1", slice=2;
        makethin, sequence=testseq;
        use, sequence=testseq;
        select, flag = error, clear;
        select, flag = error, pattern = "T1";
        ealign, dx = 0.01, dy = 0.01, arex = 0.02, arey = 0.02;
        select, flag = error, clear;
        select, flag = error, pattern = "T2";
        ealign, dx = 0.04, dy = 0.04, dpsi = 0.1;
        select, flag = error, clear;
        select, flag = error, pattern = "T3";
        ealign, dx = 0.02, dy = 0.01, arex = 0.03, arey = 0.02, dpsi = 0.1;
        select, flag = error, full;
    ''')
    line = xl.Line.from_madx_sequence(mad.sequence.testseq, install_apertures=True, apply_madx_errors=True)
    kinds = [type(e).__name__ for e in line.elements]
    assert kinds.count("XYShift") == 10 and kinds.count("SRotation") == 4 and kinds.count("LimitEllipse") == 3
    initial_x, initial_y = 0.025, -0.015
    for strict in (False, True):
        p = xl.Particles(p0c=1e9, x=initial_x, y=initial_y)
        line.track(p, strict=strict)
        assert int(p.state[0]) == 1
        assert abs(float(p.x[0]) - initial_x) < 1e-14 and abs(float(p.y[0]) - initial_y) < 1e-14


def test_error_functionality():
    """tests/test_madx_import.py:240-345: the misalignments act as intended, checked element
    by element (here: the device-side element-by-element trace of ONE launch)."""
    import xline_b200 as xl
    from xline_b200.madx_input import MadxFile

    mad = MadxFile(text='''
        T1: Collimator, L=0.0, apertype=CIRCLE, aperture={0.5};
        T2: Marker;
        T3: Collimator, L=0.0, apertype=CIRCLE, aperture={0.5};
        testseq: SEQUENCE, l = 20.0;
            T1, at =  5;
            T2, at = 10;
            T3, at = 15;
        ENDSEQUENCE;
        BEAM, PARTICLE=PROTON, ENERGY=7000.0, EXN=2.2e-6, EYN=2.2e-6;
        USE, SEQUENCE=testseq;
        select, flag = error, clear;
        select, flag = error, pattern = "T1";
        ealign, dx = 0.01, dy = 0.02, arex = 0.03, arey = 0.04;
        select, flag = error, clear;
        select, flag = error, pattern = "T3";
        ealign, dx = 0.07, dy = 0.08, dpsi = 0.7, arex = 0.08, arey = 0.09;
        select, flag = error, full;
    ''')
    line = xl.Line.from_madx_sequence(mad.sequence.testseq, install_apertures=True, apply_madx_errors=True)
    rng = np.random.default_rng(5)
    x_init, y_init = 0.1 * rng.random(10), 0.1 * rng.random(10)
    cospsi, sinpsi = math.cos(0.7), math.sin(0.7)
    for strict in (False, True):
        p = xl.Particles(p0c=1e9, x=x_init.copy(), y=y_init.copy())
        trace = line.trace_elem_by_elem(p, strict=strict).cpu().numpy()   # [len + 1, 6, n]; row i + 1 = after element i
        checked = set()
        for i, name in enumerate(line.element_names):
            x, y = trace[i + 1, 0], trace[i + 1, 2]
            if name == "t1":
                assert np.all(abs(x - (x_init - 0.01)) < 1e-14) and np.all(abs(y - (y_init - 0.02)) < 1e-14)
            elif name == "t1_aperture":
                assert np.all(abs(x - (x_init - 0.01 - 0.03)) < 1e-14)
                assert np.all(abs(y - (y_init - 0.02 - 0.04)) < 1e-14)
            elif name == "t2":
                assert np.all(abs(x - x_init) < 1e-14) and np.all(abs(y - y_init) < 1e-14)
            elif name == "t3":
                assert np.all(abs(x - (x_init - 0.07) * cospsi - (y_init - 0.08) * sinpsi) < 1e-14)
                assert np.all(abs(y + (x_init - 0.07) * sinpsi - (y_init - 0.08) * cospsi) < 1e-14)
            elif name == "t3_aperture":
                assert np.all(abs(x - (x_init - 0.07) * cospsi - (y_init - 0.08) * sinpsi - (-0.08)) < 1e-14)
                assert np.all(abs(y + (x_init - 0.07) * sinpsi - (y_init - 0.08) * cospsi - (-0.09)) < 1e-14)
            else:
                continue
            checked.add(name)
        assert checked == {"t1", "t1_aperture", "t2", "t3", "t3_aperture"}
        assert bool((p.state == 1).all())


def test_psb_c5_space_charge_apertures_monitor_against_oracle():
    """BASELINE config C5 on the real PS Booster lattice (tests/psb/*, 120 SCQGaussProfile
    kicks, 264 apertures, DipoleEdge pairs, RF, BeamMonitor), 400-particle subsample x 4 turns."""
    from xline_b200 import configs

    n, turns = 400, 4
    line, cols, p0c, m0 = configs.config_psb(n, monitor_stores=turns, monitor_ids=n)
    cols["x"][:6] *= 9.0   # make sure some particles hit the apertures
    cols["y"][6:10] *= 9.0
    ref_monitors = {}
    ref = H.run_oracle(line.to_specs(), cols, p0c, m0, num_turns=turns, monitors=ref_monitors)
    assert 0 < (ref["state"] == 0).sum() < n // 4
    # 4 turns = 480 space-charge kicks (two Faddeeva evaluations each): a few 1e-12 of the beam size
    for strict, tol in ((False, 5e-12), (True, 5e-12)):
        line.reset_monitors()
        p = make_particles(cols, p0c, m0)
        line.track(p, num_turns=turns, strict=strict)
        got = p.to_numpy()
        assert np.array_equal(got["state"], ref["state"])
        assert np.array_equal(got["at_element"], ref["at_element"])
        assert np.array_equal(got["at_turn"], ref["at_turn"])
        for k in H.COORDS:
            assert H.scaled_err(got[k], ref[k]) <= tol, (strict, k, H.scaled_err(got[k], ref[k]))
        mon = [el for el in line.elements if type(el).__name__ == "BeamMonitor"][0]
        store = list(ref_monitors.values())[0]
        written = store["at_turn"] >= 0
        for k in ("x", "px", "y", "py", "zeta", "delta"):
            g = mon.data[k].cpu().numpy()
            assert np.array_equal(np.isnan(g), ~written)
            assert H.scaled_err(g[written], store[k][written]) <= tol, (strict, k)


def _peak_tune(u):
    """Fractional tune from turn-by-turn data: Hann window + parabolic interpolation of the
    log-spectrum peak (good to ~1e-5 for 1024 turns)."""
    n = len(u)
    spec = np.abs(np.fft.rfft((u - u.mean()) * np.hanning(n)))
    k = int(np.argmax(spec[1:-1])) + 1
    a, b, c = np.log(spec[k - 1:k + 2])
    return (k + 0.5 * (a - c) / (a - 2 * b + c)) / n


def test_optics_tunes_match_tracking():
    """xline_b200.optics (first-order maps written down from the element definitions) against
    the tracker: small-amplitude turn-by-turn tunes of the bare PSB lattice, and the
    dispersion as the closed-orbit shift of an off-momentum particle."""
    import xline_b200 as xl
    from xline_b200 import configs, optics

    line, meta = configs.load_lattice("psb")
    p0c, m0 = configs.p0c_of(meta)
    tw = optics.twiss(line)
    turns = 1024
    mon = xl.BeamMonitor(num_stores=turns, start=0, skip=1, min_particle_id=0, max_particle_id=1)
    line.insert_element(0, mon, "monitor")
    d = 1e-4
    # particle 0: betatron oscillation on momentum; particle 1: on the dispersive closed orbit
    p = xl.Particles(p0c=p0c, mass0=m0, x=[1e-5, tw["dx"][0] * d], px=[0.0, tw["dpx"][0] * d],
                     y=[1e-5, 0.0], delta=[0.0, d])
    line.track(p, num_turns=turns)
    x = mon.data["x"].cpu().numpy().reshape(turns, 2)
    y = mon.data["y"].cpu().numpy().reshape(turns, 2)
    assert _peak_tune(x[:, 0]) == pytest.approx(tw["qx"] % 1, abs=2e-4)
    assert _peak_tune(y[:, 0]) == pytest.approx(tw["qy"] % 1, abs=2e-4)
    # on the closed orbit the particle stays put (up to second order in delta)
    assert np.abs(x[:, 1] - tw["dx"][0] * d).max() < 0.02 * abs(tw["dx"][0] * d)
