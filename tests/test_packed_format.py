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
            tag, size = hdr & 0xFF, (hdr >> 16) & 0xFFFF
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
