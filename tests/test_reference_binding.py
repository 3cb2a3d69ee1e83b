"""INTEGRATION.md's binding, exercised: (CPU, where /root/reference exists) the reference's OWN
element instances pack to the same lattice words as this package's classes; (GPU) the stub
``xline_b200.reference_binding.line_track`` -- NumPy arrays in a Pyparticles-like container ->
``xlb_track_host`` -> ``remove_lost_particles`` -- reproduces the golden line fixture."""
import numpy as np
import pytest

from oracle import ref_harness as rh
from tests import helpers as H


def _reference_elements(line):
    ref = rh.load_reference()
    out = []
    for name, fields in line.to_specs():
        cls = getattr(ref, name)
        allowed = set(cls().get_fields(keepextra=True)) if hasattr(cls(), "get_fields") else set(fields)
        out.append(cls(**{k: v for k, v in fields.items() if k in allowed and k != "data"}))
    return out


@pytest.mark.skipif(not rh.reference_available(), reason="reference tree absent")
@pytest.mark.parametrize("which", ["lhc", "psb", "lhc_beambeam", "fodo"])
def test_reference_element_instances_pack_to_the_same_words(which):
    """Claim of INTEGRATION.md section 1: `pack_line` needs class names and field names only, so
    the reference's own instances (xline/elements.py, be_beamfields/*.py, built by
    xline/base_classes.py:24-52) pack unchanged -- both encodings, word for word."""
    from xline_b200 import configs
    from xline_b200.lattice import pack_line

    line = {"lhc": lambda: configs.config_lhc(8)[0], "psb": lambda: configs.config_psb(8, monitor_stores=3,
                                                                                      monitor_ids=5)[0],
            "lhc_beambeam": lambda: configs.config_lhc_beambeam(8)[0], "fodo": lambda: configs.config_fodo(8)[0]}[which]()
    theirs = _reference_elements(line)
    assert all(type(e).__module__.startswith("xline.") for e in theirs)
    for strict in (False, True):
        a = pack_line(list(line.elements), strict=strict)
        b = pack_line(theirs, strict=strict)
        assert a.n_chunks == b.n_chunks and a.flags == b.flags and a.n_elements == b.n_elements
        assert np.array_equal(a.words, b.words), (which, strict)
        assert a.monitor_layout == b.monitor_layout
        assert (a.segments is None) == (b.segments is None)
        if a.segments is not None:
            assert np.array_equal(a.segments, b.segments)


class _HostParticles:
    """Pyparticles-like container on NumPy arrays (the attributes the reference's elements touch,
    SURVEY.md 8a row a2; energy bookkeeping from the oracle's restated container)."""

    def __init__(self, cols, p0c, mass0):
        from oracle import xline_oracle as xo

        n = len(cols["x"])
        o = xo.OracleParticles(n, p0c=p0c, mass0=mass0, **cols)
        for k in ("x", "px", "y", "py", "zeta", "s", "chi", "charge_ratio", "state", "particle_id",
                  "at_element", "at_turn"):
            setattr(self, k, np.array(getattr(o, k)))
        self._delta, self._rpp, self._rvv = np.array(o.delta), np.array(o.rpp), np.array(o.rvv)
        self.q0, self.mass0, self.p0c = o.q0, o.mass0, o.p0c
        self.beta0, self.gamma0, self.energy0 = o.beta0, o.gamma0, o.energy0
        self.lost_particles = []

    delta = property(lambda self: self._delta)
    rpp = property(lambda self: self._rpp)
    rvv = property(lambda self: self._rvv)

    def remove_lost_particles(self):  # tests/test_losses.py of the reference
        keep = self.state == 1
        if keep.all():
            return
        lost = {}
        for k, v in list(vars(self).items()):
            if isinstance(v, np.ndarray) and v.shape == keep.shape:
                lost[k] = v[~keep]
                setattr(self, k, v[keep])
        self.lost_particles.append(lost)


class _ForeignLine:
    """A line that is NOT this package's class: just the attribute the binding needs."""

    def __init__(self, elements):
        self.elements = elements


def _foreign_elements(specs):
    """Element instances of classes that merely carry the reference's names and fields (the GPU box
    has no /root/reference): duck typing, as for the reference's own instances."""
    out = []
    for name, fields in specs:
        cls = type(name, (), {})
        el = cls()
        for k, v in fields.items():
            setattr(el, k, v)
        if name in ("Multipole", "RFMultipole"):
            el.order = max(len(np.atleast_1d(fields.get("knl", [0]))), len(np.atleast_1d(fields.get("ksl", [0])))) - 1
        out.append(el)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("strict", [True, False])
def test_binding_stub_reproduces_the_golden_line(strict):
    from xline_b200.reference_binding import line_track
    import xline_b200 as xl

    m, specs, cols, ref = H.load_case("line_mixed_3turns")
    full_specs = []
    for name, f in specs:  # fill defaults the way the reference's dataclasses would
        d = getattr(xl, name)().to_dict(keepextra=True)
        d.pop("__class__")
        d.update(f)
        full_specs.append((name, d))
    line = _ForeignLine(_foreign_elements(full_specs))
    p = _HostParticles(cols, m["p0c"], m["mass0"])
    n = len(cols["x"])
    tally = np.zeros(len(specs), dtype=np.int64)
    for _ in range(m["num_turns"]):  # the reference's callers write the turn loop themselves
        line_track(line, p, strict=strict, loss_tally=tally)
    lost_ids = np.concatenate([lp["particle_id"] for lp in p.lost_particles]) if p.lost_particles else np.zeros(0, int)
    assert len(p.x) + len(lost_ids) == n and len(lost_ids) == int((ref["state"] == 0).sum()) > 0
    assert int(tally.sum()) == len(lost_ids)
    # survivors, in their original order (compaction keeps it)
    alive = ref["state"] == 1
    assert np.array_equal(p.particle_id, np.flatnonzero(alive))
    tol = 1e-14 if strict else 1e-12
    for k in H.COORDS:
        got = getattr(p, k)
        assert H.scaled_err(got, ref[k][alive]) <= tol, (k, H.scaled_err(got, ref[k][alive]))
    assert np.array_equal(p.at_turn, ref["at_turn"][alive])
    # the lost ones: frozen at their aperture, element and turn recorded
    got_elem = np.concatenate([lp["at_element"] for lp in p.lost_particles])
    got_turn = np.concatenate([lp["at_turn"] for lp in p.lost_particles])
    order = np.argsort(lost_ids)
    assert np.array_equal(lost_ids[order], np.flatnonzero(~alive))
    assert np.array_equal(got_elem[order], ref["at_element"][~alive])
    assert np.array_equal(got_turn[order], ref["at_turn"][~alive])
    for k in ("x", "y"):
        got = np.concatenate([lp[k] for lp in p.lost_particles])[order]
        assert H.scaled_err(got, ref[k][~alive]) <= tol, k
