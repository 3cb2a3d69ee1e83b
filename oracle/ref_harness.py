"""TEST INFRASTRUCTURE ONLY (container-side).  Executes the *real* reference code.

Imports the reference's element modules straight from ``/root/reference`` with the
package ``__init__`` bypassed (``xline/__init__.py:3`` raises unconditionally) and
drives them with a particle stand-in.  This module cannot travel to the GPU box
(``/root/reference`` does not exist there); it is used only

* by ``oracle/make_golden.py`` to generate the committed fixtures in ``tests/golden``;
* by ``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent) to pin
  the numpy restatement ``oracle/xline_oracle.py`` against the reference's own code.

What is NOT under ``/root/reference`` and is therefore restated here (recalled from the
public ``pysixtrack``/``xpart`` ``Pyparticles`` source, *unpinned* – see SURVEY.md §8c):
the particle container (``xline/particles.py:1-6`` subclasses
``xpart.particles._pyparticles.Pyparticles``) and the 7-line ``Line.track`` loop
(``xline/line.py:89-95``; ``xline/line.py`` itself imports ``xpart``/``xobjects``).
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("XLINE_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "xline", "elements.py"))


_loaded = {}


def load_reference():
    """Return the reference's ``xline.elements`` module (import with __init__ bypassed)."""
    if "elements" in _loaded:
        return _loaded["elements"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not hasattr(np, "float_"):  # propagate_sigma_matrix.py:5 uses the NumPy-1 alias
        np.float_ = np.float64
    pkg = types.ModuleType("xline")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "xline")]
    saved = sys.modules.get("xline")
    sys.modules["xline"] = pkg
    try:
        import importlib

        elements = importlib.import_module("xline.elements")
        mathlibs = importlib.import_module("xline.mathlibs")
        importlib.import_module("xline.be_beamfields.slicing")
        importlib.import_module("xline.loader_sixtrack")
    finally:
        # leave the submodules registered (they reference each other) but make sure a
        # later ``import xline`` of anything else cannot pick up the stub silently.
        if saved is not None:
            sys.modules["xline"] = saved
    _loaded["elements"] = elements
    _loaded["mathlibs"] = mathlibs
    _loaded["pkg"] = pkg
    return elements


def reference_module(name):
    load_reference()
    return sys.modules["xline." + name]


class RefParticles:
    """Stand-in for ``xpart.particles._pyparticles.Pyparticles`` (restated, unpinned).

    Only what the reference's elements touch: attributes ``x px y py zeta delta s rpp rvv
    chi charge_ratio q0 p0c beta0 mass0 state _m`` and the methods ``add_to_energy``
    (``elements.py:227,245``), ``remove_lost_particles`` (``elements.py:420,442,474``),
    ``copy``, and the ``delta`` setter (``beambeam.py:280-283``).
    """

    _array_fields = (
        "x px y py zeta s _delta _rpp _rvv chi charge_ratio state particle_id "
        "at_element at_turn"
    ).split()

    def __init__(self, p0c=1e9, mass0=938.27208816e6, q0=1.0, n=None, **kw):
        load_reference()
        self._m = _loaded["mathlibs"].MathlibDefault
        self.q0 = float(q0)
        self.mass0 = float(mass0)
        self.p0c = float(p0c)
        self.energy0 = float(np.sqrt(self.p0c ** 2 + self.mass0 ** 2))
        self.beta0 = self.p0c / self.energy0
        self.gamma0 = self.energy0 / self.mass0
        scalar = n is None and not any(
            hasattr(v, "__len__") for v in kw.values()
        )
        if n is None and not scalar:
            n = max(len(v) for v in kw.values() if hasattr(v, "__len__"))
        self._scalar = scalar

        def mk(val, dtype=np.float64):
            if scalar:
                return dtype(val)
            return np.array(np.broadcast_to(np.asarray(val, dtype=dtype), (n,)))

        for name in "x px y py zeta s".split():
            setattr(self, name, mk(kw.get(name, 0.0)))
        self.chi = mk(kw.get("chi", 1.0))
        self.charge_ratio = mk(kw.get("charge_ratio", 1.0))
        self.state = mk(kw.get("state", 1), np.int64)
        self.particle_id = (
            np.int64(0) if scalar else np.arange(n, dtype=np.int64)
        )
        self.at_element = mk(0, np.int64)
        self.at_turn = mk(0, np.int64)
        self.delta = mk(kw.get("delta", 0.0))
        self.lost_particles = []

    # -- energy bookkeeping (recalled Pyparticles arithmetic) -------------------------
    @property
    def delta(self):
        return self._delta

    @delta.setter
    def delta(self, delta):
        sqrt = np.sqrt
        self._delta = delta
        deltabeta0 = delta * self.beta0
        ptaubeta0 = sqrt(deltabeta0 ** 2 + 2 * deltabeta0 * self.beta0 + 1) - 1
        one_plus_delta = 1 + delta
        self._rvv = one_plus_delta / (1 + ptaubeta0)
        self._rpp = 1 / one_plus_delta

    @property
    def rpp(self):
        return self._rpp

    @property
    def rvv(self):
        return self._rvv

    def add_to_energy(self, energy):
        sqrt = np.sqrt
        oldrvv = self._rvv
        deltabeta0 = self._delta * self.beta0
        ptaubeta0 = sqrt(deltabeta0 ** 2 + 2 * deltabeta0 * self.beta0 + 1) - 1
        ptaubeta0 = ptaubeta0 + energy / self.energy0
        ptau = ptaubeta0 / self.beta0
        self._delta = sqrt(ptau ** 2 + 2 * ptau / self.beta0 + 1) - 1
        one_plus_delta = 1 + self._delta
        self._rvv = one_plus_delta / (1 + ptaubeta0)
        self._rpp = 1 / one_plus_delta
        self.zeta = self.zeta * (self._rvv / oldrvv)

    # -- loss compaction -----------------------------------------------------------------
    def remove_lost_particles(self, keep_memory=True):
        if self._scalar:
            return
        keep = self.state == 1
        if np.all(keep):
            return
        lost = ~keep
        if keep_memory:
            rec = {f: getattr(self, f)[lost].copy() for f in self._array_fields}
            self.lost_particles.append(rec)
        for f in self._array_fields:
            setattr(self, f, getattr(self, f)[keep])

    def copy(self):
        import copy as _copy

        new = _copy.copy(self)
        for f in self._array_fields:
            v = getattr(self, f)
            setattr(new, f, v.copy() if hasattr(v, "copy") else v)
        new.lost_particles = list(self.lost_particles)
        return new

    def __len__(self):
        return 1 if self._scalar else len(self.x)


def ref_line_track(elements, p, num_turns=1, monitor_hook=None):
    """Restated ``Line.track`` (``xline/line.py:89-95``) wrapped in a turn loop that also
    derives ``at_element``/``at_turn`` for lost particles (the reference never stores
    them: a particle that disappears from the arrays in element ``i`` during turn ``t``
    gets ``at_element=i, at_turn=t``; survivors end with ``at_turn=t0+num_turns``).
    Returns a dict of full-length arrays indexed by ``particle_id``."""
    n0 = len(p)
    ids0 = np.array(p.particle_id, copy=True)
    for _ in range(num_turns):
        for iel, el in enumerate(elements):
            before = len(p.lost_particles)
            if type(el).__name__ == "BeamMonitor":
                if monitor_hook is not None:
                    monitor_hook(iel, el, p)
                continue
            el.track(p)
            if len(p.lost_particles) > before:
                rec = p.lost_particles[-1]
                rec["at_element"][:] = iel
                rec["state"][:] = 0
            if len(p) == 0:
                break
        p.at_turn = p.at_turn + 1
        p.at_element = p.at_element * 0
        if len(p) == 0:
            break
    out = {}
    fields = "x px y py zeta s _delta _rpp _rvv state at_element at_turn".split()
    pos = {int(pid): i for i, pid in enumerate(ids0)}
    for f in fields:
        proto = getattr(p, f)
        arr = np.zeros(n0, dtype=proto.dtype)
        idx = np.array([pos[int(i)] for i in p.particle_id], dtype=np.int64)
        arr[idx] = proto
        for rec in p.lost_particles:
            ridx = np.array([pos[int(i)] for i in rec["particle_id"]], dtype=np.int64)
            arr[ridx] = rec[f]
        out[f.lstrip("_")] = arr
    out["particle_id"] = ids0
    return out
