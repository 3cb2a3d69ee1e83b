"""The binding a maintainer of the reference would add: ``Line.track`` of the reference
(``xline/line.py:89-95``: ``for el in self.elements: el.track(p)``) served by ``xlb_track_host``.

``line_track(line, p, num_turns=1)`` takes the reference's OWN objects -- a ``Line`` (anything with
an ``elements`` list of the reference's element instances, ``xline/elements.py``; only class names
and ``_description`` field names are used, ``xline/base_classes.py:24-52``) and a particle
container with the attributes of xpart's ``Pyparticles`` (``xline/particles.py:1-6``) holding NumPy
arrays -- packs the line once, hands host pointers to the C ABI (``include/xline_b200.h``), and
applies the reference's loss semantics (``p.remove_lost_particles()``, ``xline/elements.py:420``).
Installing it in the reference is one line::

    from xline_b200.reference_binding import line_track
    xline.Line.track = line_track

This is INTEGRATION.md section 2 as code; ``tests/test_reference_binding.py`` runs it.
"""
import ctypes as C

import numpy as np

from . import _cabi
from .lattice import pack_line

_F64 = ("x", "px", "y", "py", "zeta", "s")
_I64 = ("state", "at_element", "at_turn", "particle_id")


def _packed_for(line, strict):
    cache = line.__dict__.setdefault("_xlb_packed", {})
    key = (bool(strict), tuple(map(id, line.elements)))
    hit = cache.get(bool(strict))
    if hit is None or hit[0] != key:
        packed = pack_line(list(line.elements), strict=strict)
        lat = packed.c_lattice()
        _cabi.check(_cabi.lib().xlb_lattice_validate(C.byref(lat)))
        hit = cache[bool(strict)] = (key, packed)
    return hit[1]


def line_track(line, p, num_turns=1, strict=False, turns_per_launch=0, loss_tally=None):
    """In place on ``p`` like the reference's ``Line.track``; returns ``None``.  Elements edited
    in place after the first call need ``line.__dict__.pop("_xlb_packed")`` (the reference's
    classes carry no edit tracking)."""
    packed = _packed_for(line, strict)
    lat = packed.c_lattice()  # xlb_lattice_t incl. the segment table
    n = int(np.size(p.x))
    if n == 0 or num_turns == 0:
        return None
    cols = {k: np.ascontiguousarray(np.broadcast_to(getattr(p, k), (n,)), dtype=np.float64).copy() for k in _F64}
    for k in ("delta", "rpp", "rvv", "chi", "charge_ratio"):
        cols[k] = np.ascontiguousarray(np.broadcast_to(getattr(p, k), (n,)), dtype=np.float64).copy()
    defaults = {"state": 1, "at_element": 0, "at_turn": 0, "particle_id": np.arange(n)}
    for k in _I64:
        cols[k] = np.ascontiguousarray(np.broadcast_to(getattr(p, k, defaults[k]), (n,)), dtype=np.int64).copy()
    cp = _cabi.Particles()
    cp.n = n
    for k, a in cols.items():
        setattr(cp, k, a.ctypes.data)
    cp.q0, cp.mass0, cp.p0c = float(p.q0), float(p.mass0), float(p.p0c)
    cp.beta0, cp.energy0 = float(p.beta0), float(p.energy0)
    cp.gamma0 = float(getattr(p, "gamma0", p.energy0 / p.mass0))
    opts = _cabi.TrackOptions()
    opts.num_turns = int(num_turns)
    opts.turns_per_launch = int(turns_per_launch)
    if loss_tally is not None:
        assert loss_tally.dtype == np.int64 and loss_tally.size >= packed.n_elements and loss_tally.flags.c_contiguous
        opts.loss_tally = loss_tally.ctypes.data
    mon = None
    if packed.monitor_words > 0:
        mon = line.__dict__.get("_xlb_monitor")
        if mon is None or mon.size != packed.monitor_words:
            mon = line.__dict__["_xlb_monitor"] = np.full(packed.monitor_words, np.nan)
        opts.monitor_data = mon.ctypes.data
        opts.monitor_words = packed.monitor_words
    _cabi.check(_cabi.lib().xlb_track_host(C.byref(lat), C.byref(cp), C.byref(opts)))
    for k in _F64 + _I64:  # results are in the host arrays
        setattr(p, k, cols[k])
    if hasattr(p, "_delta"):  # Pyparticles keeps delta / rpp / rvv consistent behind a property
        p._delta, p._rpp, p._rvv = cols["delta"], cols["rpp"], cols["rvv"]
    else:
        p.delta, p.rpp, p.rvv = cols["delta"], cols["rpp"], cols["rvv"]
    if mon is not None:
        from .lattice import MONITOR_FIELDS

        for slot in packed.monitor_layout:
            ns, nn = slot["num_stores"], slot["nn"]
            if ns > 0 and nn > 0:
                view = mon[slot["offset"]: slot["offset"] + len(MONITOR_FIELDS) * ns * nn].reshape(
                    len(MONITOR_FIELDS), ns, nn)
                line.elements[slot["element_index"]].data = {k: view[i] for i, k in enumerate(MONITOR_FIELDS)}
    if hasattr(p, "remove_lost_particles"):
        p.remove_lost_particles()  # reference semantics: lost particles leave the arrays
    return None
