"""The packed lattice format and the packer, checked without a GPU: the words written by
``xline_b200.lattice.pack_line`` are interpreted by ``tests/packed_interpreter.py`` strictly as
``include/xline_b200.h`` documents them and the result is compared with the oracle.

* strict encoding (raw parameters, reference operation order; fused records included):
  **bit-identical** to the oracle -- NumPy evaluates without FMA contraction, like the strict
  kernel;
* fast encoding (``knl/i!`` folded, reciprocals, merged co-located multipoles): equal to
  rounding, same losses at the same elements and turns.
"""
import numpy as np
import pytest

from tests import helpers as H
from tests import packed_interpreter as PI
from xline_b200 import configs


def _subset(cols, n, boost=1.0):
    out = {k: np.ascontiguousarray(v[:n]) for k, v in cols.items() if k != "particle_id"}
    for k in ("x", "px", "y", "py"):
        out[k] = out[k] * boost
    return out


def _cases():
    import xline_b200 as xl

    line, cols, p0c, m0 = configs.config_fodo(400)
    fodo = xl.Line(list(line.elements) + [
        xl.LimitRect(min_x=-4e-3, max_x=4e-3, min_y=-4e-3, max_y=4e-3),
        xl.SRotation(angle=7.0), xl.XYShift(dx=1e-4, dy=-2e-4),
        xl.LimitRectEllipse(max_x=5e-3, max_y=4e-3, a=6e-3, b=4.5e-3),
        xl.XYShift(dx=-1e-4, dy=2e-4), xl.SRotation(angle=-7.0),
        xl.LimitRect(min_x=-3e-3, max_x=5e-3, min_y=-4e-3, max_y=3.5e-3),  # not symmetric
    ])
    yield "fodo", fodo, _subset(cols, 400), p0c, m0, 3
    line, cols, p0c, m0 = configs.config_lhc(4000)
    order = np.argsort(-np.hypot(cols["x"], cols["y"]))  # large amplitudes first: some get lost
    big = {k: v[order] for k, v in cols.items()}
    yield "lhc", line, _subset(big, 48, boost=1.6), p0c, m0, 1
    line, cols, p0c, m0 = configs.config_petra4(64)
    yield "petra4", line, _subset(cols, 64), p0c, m0, 1


CASES = {c[0]: c[1:] for c in _cases()}


@pytest.mark.parametrize("name", sorted(CASES))
def test_strict_encoding_interpreted_is_the_oracle_bit_for_bit(name):
    line, cols, p0c, m0, turns = CASES[name]
    packed = line.pack(strict=True)
    assert packed.strict
    got = PI.track(packed, cols, p0c, m0, num_turns=turns)
    ref = H.run_oracle(line.to_specs(), cols, p0c, m0, num_turns=turns)
    if name != "petra4":
        assert (ref["state"] == 0).any(), "case should exercise the loss bookkeeping"
    for k in ("state", "at_element", "at_turn"):
        assert np.array_equal(got[k], ref[k]), k
    for k in H.COORDS + ("rpp", "rvv"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k


@pytest.mark.parametrize("name", sorted(CASES))
def test_fast_encoding_interpreted_matches_the_oracle(name):
    line, cols, p0c, m0, turns = CASES[name]
    packed = line.pack(strict=False)
    got = PI.track(packed, cols, p0c, m0, num_turns=turns)
    ref = H.run_oracle(line.to_specs(), cols, p0c, m0, num_turns=turns)
    for k in ("state", "at_element", "at_turn"):
        assert np.array_equal(got[k], ref[k]), k
    alive = ref["state"] == 1
    for k in H.COORDS:
        assert H.scaled_err(got[k][alive], ref[k][alive]) <= 1e-10, (k, H.scaled_err(got[k][alive], ref[k][alive]))
        # frozen at the aperture: same map up to there, same tolerance
        if (~alive).any():
            assert H.scaled_err(got[k][~alive], ref[k][~alive]) <= 1e-9, k


def test_fast_lhc_lattice_uses_every_block_family():
    """The interpreter is only a check of the format if the lattice exercises it: the fast C2
    lattice must contain thin blocks of several aperture kinds, curved ones and merged ones."""
    line = CASES["lhc"][0]
    words = np.asarray(line.pack(strict=False).words, dtype=np.uint64)
    tags = set()
    packed = line.pack(strict=False)
    for ch in range(packed.n_chunks):
        w = ch * packed.chunk_words
        while True:
            hdr = int(words[w])
            tag, size = hdr & 0xFF, (hdr >> 16) & 0x3FFF
            if tag in (PI.T_END_CHUNK, PI.T_END_TURN):
                break
            tags.add(tag)
            w += 2 * size
    assert {0x8D, 0x88, 0x89, 0x8B, 0xA9} <= tags, sorted(hex(t) for t in tags)


def test_rfmultipole_and_monitor_records():
    """RFMultipole (strict: the oracle bit for bit) and the BeamMonitor record: slot arithmetic
    of elements.py:497-524 and the [7 fields][num_stores][n_ids] storage layout, with skip,
    a particle-id window, losses in between and a rolling monitor."""
    import xline_b200 as xl

    n, turns = 300, 9
    rng = np.random.default_rng(21)
    cols = dict(x=rng.normal(0, 1e-3, n), px=rng.normal(0, 1e-4, n), y=rng.normal(0, 1e-3, n),
                py=rng.normal(0, 1e-4, n), zeta=rng.normal(0, 0.05, n), delta=rng.normal(0, 3e-4, n))
    mon = xl.BeamMonitor(num_stores=3, start=2, skip=2, min_particle_id=10, max_particle_id=259)
    roll = xl.BeamMonitor(num_stores=2, start=0, skip=1, min_particle_id=0, max_particle_id=n - 1, is_rolling=True)
    line = xl.Line([
        xl.Drift(length=1.0), xl.Multipole(knl=[0, 0.3]), xl.LimitEllipse(a=3e-3, b=3e-3), mon,
        xl.RFMultipole(voltage=2e5, frequency=4e8, lag=30.0, knl=[1e-4, 0.02, 0.5], ksl=[0, 0.01],
                       pn=[10.0, 20.0, 0.0], ps=[0.0, 45.0]),
        xl.Drift(length=2.0), xl.Multipole(knl=[0, -0.3]), roll, xl.Cavity(voltage=1e6, frequency=4e8, lag=180),
    ])
    p0c, m0 = 450e9, 938.27208816e6
    packed = line.pack(strict=True)
    buf = np.full(packed.monitor_words, np.nan)
    got = PI.track(packed, cols, p0c, m0, num_turns=turns, monitor=buf)
    stores = {}
    ref = H.run_oracle(line.to_specs(), cols, p0c, m0, num_turns=turns, monitors=stores)
    assert (ref["state"] == 0).any()
    for k in ("state", "at_element", "at_turn"):
        assert np.array_equal(got[k], ref[k]), k
    for k in H.COORDS:
        assert np.array_equal(got[k], ref[k]), k
    for slot in packed.monitor_layout:
        want = stores[slot["element_index"]]
        ns, nn = slot["num_stores"], slot["nn"]
        view = buf[slot["offset"]: slot["offset"] + 7 * ns * nn].reshape(7, ns, nn)
        written = want["at_turn"] >= 0
        assert written.any() and not written.all()
        for f, k in enumerate(("x", "px", "y", "py", "zeta", "delta")):
            assert np.array_equal(np.isnan(view[f]), ~written), k
            assert np.array_equal(view[f][written], want[k][written]), k
        assert np.array_equal(view[6][written], want["at_turn"][written].astype(float))


def _random_line(rng):
    """A random thin-lens line, biased towards the sequences the packer's peepholes look for
    (multipole -> aperture -> drift, co-located multipoles, dipole edge -> drift, aperture -> drift)
    and towards
    their edge cases (exact no-ops, zero-length curved multipoles, asymmetric boxes)."""
    import xline_b200 as xl

    def multipole():
        order = int(rng.integers(0, 7))
        # strengths that keep a millimetre beam physical over three turns: a beam blown up to
        # px^2 + py^2 > (1 + delta)^2 turns into NaN at a zero-length DriftExact in the reference,
        # which the packer drops as an exact no-op
        knl = (rng.normal(0, 1, order + 1) * 0.05 * 3.0 ** np.arange(order + 1)).tolist()
        ksl = (rng.normal(0, 1, order + 1) * 0.05 * 3.0 ** np.arange(order + 1)).tolist()
        kind = rng.integers(0, 8)
        if kind == 0:
            return xl.Multipole(knl=[0.0] * (order + 1), ksl=[0.0] * (order + 1))  # exact no-op
        if kind == 1:
            ksl = [0.0]
        if kind in (2, 3):  # curved; length 0 switches the hxx/hyy terms off (elements.py:141-147)
            return xl.Multipole(knl=knl, ksl=ksl, hxl=rng.normal(0, 1e-3), hyl=rng.normal(0, 1e-4) * (kind == 3),
                                length=float(rng.choice([0.0, 0.7])))
        return xl.Multipole(knl=knl, ksl=ksl)

    def aperture():
        k = rng.integers(0, 4)
        if k == 0:
            return xl.LimitRect(min_x=-3e-3, max_x=3e-3, min_y=-2.5e-3, max_y=2.5e-3)
        if k == 1:
            return xl.LimitRect(min_x=-2e-3, max_x=3.5e-3, min_y=-3e-3, max_y=2e-3)
        if k == 2:
            return xl.LimitEllipse(a=3.2e-3, b=2.6e-3)
        return xl.LimitRectEllipse(max_x=3e-3, max_y=2.5e-3, a=3.5e-3, b=3.1e-3)

    def drift():
        L = float(rng.choice([0.0, 0.3, 1.1, 2.5]))
        return xl.DriftExact(length=L) if rng.random() < 0.3 else xl.Drift(length=L)

    els = []
    for _ in range(int(rng.integers(6, 30))):
        u = rng.random()
        if u < 0.45:
            els.append(multipole())
            if rng.random() < 0.5:
                els.append(aperture())
            if rng.random() < 0.4:
                els.append(multipole())
                if rng.random() < 0.5:
                    els.append(aperture())
            if rng.random() < 0.8:
                els.append(drift())
        elif u < 0.6:
            if rng.random() < 0.35:  # a collimator between two drifts: aperture -> drift in one record
                els.append(aperture())
            els.append(drift())
        elif u < 0.7:
            els.append(xl.DipoleEdge(h=0.01, e1=rng.normal(0, 0.05), hgap=0.02, fint=0.5))
            if rng.random() < 0.7:
                els.append(drift())
        elif u < 0.8:
            els.append(xl.SRotation(angle=rng.normal(0, 5)))
        elif u < 0.9:
            els.append(xl.XYShift(dx=rng.normal(0, 1e-4), dy=rng.normal(0, 1e-4)))
        else:
            els.append(xl.Cavity(voltage=1e6, frequency=4e8, lag=float(rng.choice([0.0, 180.0]))))
    return xl.Line(els)


@pytest.mark.parametrize("seed", range(40))
def test_random_lines_strict_bitwise_and_fast_close(seed):
    """Fuzz of the packer: random lines, small chunks (records get cut across many chunks),
    three turns with losses.  Strict encoding == oracle bit for bit; fast encoding (folding,
    fusing, merging) to rounding with identical loss bookkeeping."""
    rng = np.random.default_rng(1000 + seed)
    line = _random_line(rng)
    line.chunk_words = int(rng.choice([32, 64, 256]))
    n = 150
    cols = dict(x=rng.normal(0, 8e-4, n), px=rng.normal(0, 1e-4, n), y=rng.normal(0, 8e-4, n),
                py=rng.normal(0, 1e-4, n), zeta=rng.normal(0, 0.05, n), delta=rng.normal(0, 3e-4, n))
    p0c, m0 = 26e9, 938.27208816e6
    with np.errstate(all="ignore"):
        ref = H.run_oracle(line.to_specs(), cols, p0c, m0, num_turns=3)
        strict = PI.track(line.pack(strict=True), cols, p0c, m0, num_turns=3)
        fast = PI.track(line.pack(strict=False), cols, p0c, m0, num_turns=3)
    for k in ("state", "at_element", "at_turn"):
        assert np.array_equal(strict[k], ref[k]), k
    for k in H.COORDS:
        assert np.array_equal(strict[k], ref[k], equal_nan=True), k
    # fast: a particle within rounding of an aperture edge may legitimately differ; none does here
    for k in ("state", "at_element", "at_turn"):
        assert np.array_equal(fast[k], ref[k]), k
    for k in H.COORDS:
        assert H.scaled_err(fast[k], ref[k]) <= 1e-10, (k, H.scaled_err(fast[k], ref[k]))


def test_beam_field_records():
    """BeamBeam4D and the three space-charge elements (round and both elliptical orientations,
    q-Gaussian profiles with q = 1, > 1 and < 1, linear and cubic-spline line densities): record
    layouts and the constants the fast encoding folds at pack time, against the oracle."""
    import xline_b200 as xl

    n = 400
    rng = np.random.default_rng(33)
    cols = dict(x=rng.normal(0, 2e-3, n), px=rng.normal(0, 1e-4, n), y=rng.normal(0, 2e-3, n),
                py=rng.normal(0, 1e-4, n), zeta=rng.normal(0, 0.4, n), delta=rng.normal(0, 1e-3, n))
    cols["x"][:3] = [0.0, 1e-12, -1e-3]  # on axis / inside the linearised core / negative quadrant
    cols["y"][:3] = [0.0, 0.0, -2e-3]
    prof = np.exp(-0.5 * (np.linspace(-1.5, 1.5, 31) / 0.5) ** 2)
    line = xl.Line([
        xl.BeamBeam4D(charge=1e11, sigma_x=1.2e-3, sigma_y=0.7e-3, beta_r=0.9, x_bb=1e-4, y_bb=-2e-4, d_px=1e-7, d_py=-2e-7),
        xl.Drift(length=1.0),
        xl.BeamBeam4D(charge=-8e10, sigma_x=0.6e-3, sigma_y=1.5e-3, beta_r=1.0),
        xl.BeamBeam4D(charge=5e10, sigma_x=1.0e-3, sigma_y=1.0e-3 * (1 + 1e-12), beta_r=1.0),  # round branch
        xl.Multipole(knl=[0, 0.1]), xl.Drift(length=0.5),
        xl.SCCoasting(number_of_particles=3e12, circumference=157.0, sigma_x=2e-3, sigma_y=1e-3, length=3.0,
                      x_co=1e-4, y_co=2e-4),
        xl.SCQGaussProfile(number_of_particles=1e11, bunchlength_rms=0.4, sigma_x=1e-3, sigma_y=2.5e-3, length=2.0,
                           q_parameter=1.0),
        xl.SCQGaussProfile(number_of_particles=1e11, bunchlength_rms=0.4, sigma_x=1.5e-3, sigma_y=1.5e-3, length=2.0,
                           q_parameter=1.3),
        xl.SCQGaussProfile(number_of_particles=1e11, bunchlength_rms=0.4, sigma_x=2e-3, sigma_y=1.5e-3, length=2.0,
                           q_parameter=0.7),
        xl.Drift(length=0.7),
        xl.SCInterpolatedProfile(number_of_particles=2e11, line_density_profile=prof.tolist(), dz=0.1, z0=-1.5,
                                 sigma_x=1.1e-3, sigma_y=0.9e-3, length=1.5, method=0),
        xl.SCInterpolatedProfile(number_of_particles=2e11, line_density_profile=prof.tolist(), dz=0.1, z0=-1.5,
                                 sigma_x=0.9e-3, sigma_y=1.4e-3, length=1.5, method=1),
        xl.LimitEllipse(a=8e-3, b=8e-3),
    ])
    p0c, m0 = 0.571e9, 938.27208816e6  # PS Booster momentum: space charge matters there
    with np.errstate(all="ignore"):
        ref = H.run_oracle(line.to_specs(), cols, p0c, m0, num_turns=2)
    kick = np.sqrt(np.mean((ref["px"] - cols["px"]) ** 2))
    assert kick > 1e-7, "the lattice must actually kick"
    for strict in (True, False):
        packed = line.pack(strict=strict)
        assert packed.flags & 2  # XLB_F_BEAMFIELDS
        with np.errstate(all="ignore"):
            got = PI.track(packed, cols, p0c, m0, num_turns=2)
        for k in ("state", "at_element", "at_turn"):
            assert np.array_equal(got[k], ref[k]), (k, strict)
        for k in H.COORDS:
            assert H.scaled_err(got[k], ref[k]) <= 1e-12, (k, strict, H.scaled_err(got[k], ref[k]))


def test_psb_c5_lattice_interpreted():
    """The real C5 lattice (PS Booster: 120 space-charge kicks, 264 apertures of three kinds,
    RF, one BeamMonitor) through the format interpreter, two turns, with particles pushed into
    the apertures: both encodings against the oracle, monitor contents included."""
    n, turns = 96, 2
    line, cols, p0c, m0 = configs.config_psb(n, monitor_stores=turns, monitor_ids=n)
    cols = {k: np.array(v) for k, v in cols.items() if k != "particle_id"}
    cols["x"][::7] *= 12.0  # some beyond the vacuum chamber
    cols["y"][3::11] *= 15.0
    stores = {}
    with np.errstate(all="ignore"):
        ref = H.run_oracle(line.to_specs(), cols, p0c, m0, num_turns=turns, monitors=stores)
    assert 0 < (ref["state"] == 0).sum() < n
    for strict in (True, False):
        packed = line.pack(strict=strict)
        buf = np.full(packed.monitor_words, np.nan)
        with np.errstate(all="ignore"):
            got = PI.track(packed, cols, p0c, m0, num_turns=turns, monitor=buf)
        for k in ("state", "at_element", "at_turn"):
            assert np.array_equal(got[k], ref[k]), (k, strict)
        for k in H.COORDS:
            assert H.scaled_err(got[k], ref[k]) <= 1e-11, (k, strict, H.scaled_err(got[k], ref[k]))
        (slot,) = packed.monitor_layout
        want = stores[slot["element_index"]]
        view = buf[slot["offset"]: slot["offset"] + 7 * slot["num_stores"] * slot["nn"]].reshape(7, slot["num_stores"], slot["nn"])
        written = want["at_turn"] >= 0
        assert np.array_equal(np.isnan(view[0]), ~written)
        for f, k in enumerate(("x", "px", "y", "py", "zeta", "delta")):
            assert H.scaled_err(view[f][written], want[k][written]) <= 1e-11, (k, strict)
