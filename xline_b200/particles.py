"""Particle container of the drop-in API: structure-of-arrays torch tensors (fp64 /
int64) resident in HBM, with the attribute names and energy bookkeeping of the
reference's particle class (``xline/particles.py:1-6`` -> xpart ``Pyparticles``; the
arithmetic is restated from the public pysixtrack/xpart source).

Lost particles are not removed by tracking: they stay in the arrays with ``state == 0``,
``at_element`` / ``at_turn`` set and coordinates frozen at the aperture.
``remove_lost_particles()`` gives the reference's compacted view (``tests/test_losses.py``)
and appends the removed ones to ``lost_particles``.
"""
import math

import numpy as np
import torch

PROTON_MASS_EV = 938.27208816e6

FLOAT_COLS = ("x", "px", "y", "py", "zeta", "delta", "rpp", "rvv", "s", "chi", "charge_ratio")
INT_COLS = ("state", "at_element", "at_turn", "particle_id")


def _default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


class MathlibDefault:
    """The reference's math seam (``xline/mathlibs.py:5-18``): elements reach NumPy/SciPy only
    through ``p._m``.  Irrelevant on the device; kept on the host container so code written
    against ``p._m`` keeps working."""

    from numpy import sqrt, exp, sin, cos, abs, pi, tan, interp, linspace  # noqa: A004
    from numpy import power as pow  # noqa: A004

    @classmethod
    def wfun(cls, z_re, z_im):
        from scipy.special import wofz

        w = wofz(z_re + 1j * z_im)
        return w.real, w.imag

    @classmethod
    def gamma(cls, arg):
        from scipy.special import gamma as tgamma

        assert arg > 0.0
        return tgamma(arg)


class Particles:
    """``Particles(p0c=..., x=..., px=..., ...)``; scalars broadcast to the longest array.

    Reference quantities (``q0 mass0 p0c beta0 gamma0 energy0``) are scalars shared by the
    whole set, as in every reference example.  ``device`` defaults to the current CUDA
    device; a CPU-resident container can be built for host-side work (packing, sharding,
    I/O) but ``Line.track`` refuses it -- there is no CPU tracking path in this package.
    """

    _m = MathlibDefault

    def __init__(self, p0c=1e9, mass0=PROTON_MASS_EV, q0=1.0, n=None, device=None, pinned=False,
                 **cols):
        self.device = torch.device(device) if device is not None else _default_device()
        self.q0 = float(q0)
        self.mass0 = float(mass0)
        self._set_p0c(float(p0c))
        lens = [len(v) for v in cols.values() if hasattr(v, "__len__")]
        if n is None:
            n = max(lens) if lens else 1
        self._scalar_input = not lens and "n" not in cols
        self._pinned = bool(pinned) and self.device.type == "cpu"
        unknown = set(cols) - set(FLOAT_COLS) - set(INT_COLS)
        if unknown:
            raise TypeError("unknown particle attribute(s): %s" % sorted(unknown))

        def mk(val, dtype):
            t = torch.as_tensor(np.asarray(val), dtype=dtype)
            t = t.expand(n).clone() if t.ndim == 0 else t.clone()
            if t.shape != (n,):
                raise ValueError("particle column has length %d, expected %d" % (t.numel(), n))
            if self._pinned:
                t = t.pin_memory()
            return t.to(self.device)

        for k in ("x", "px", "y", "py", "zeta", "s"):
            setattr(self, k, mk(cols.get(k, 0.0), torch.float64))
        self.chi = mk(cols.get("chi", 1.0), torch.float64)
        self.charge_ratio = mk(cols.get("charge_ratio", 1.0), torch.float64)
        self.state = mk(cols.get("state", 1), torch.int64)
        self.at_element = mk(cols.get("at_element", 0), torch.int64)
        self.at_turn = mk(cols.get("at_turn", 0), torch.int64)
        pid = cols.get("particle_id")
        self.particle_id = mk(pid if pid is not None else np.arange(n), torch.int64)
        # delta -> (rpp, rvv) in NumPy on the host: the same IEEE operations, in the same
        # order, as the reference's Pyparticles setter, independent of the device the
        # columns end up on (torch's CPU and CUDA element-wise kernels differ in the last bit)
        d = np.array(np.broadcast_to(np.asarray(cols.get("delta", 0.0), dtype=np.float64), (n,)))
        b0 = self._beta0
        db0 = d * b0
        ptaub0 = np.sqrt(db0 ** 2 + 2 * db0 * b0 + 1) - 1
        opd = 1 + d
        self._delta = mk(d, torch.float64)
        self.rvv = mk(opd / (1 + ptaub0), torch.float64)
        self.rpp = mk(1 / opd, torch.float64)
        if "rpp" in cols:
            self.rpp = mk(cols["rpp"], torch.float64)
        if "rvv" in cols:
            self.rvv = mk(cols["rvv"], torch.float64)
        self.lost_particles = []

    # -- reference particle ------------------------------------------------------------
    def _set_p0c(self, p0c):
        self._p0c = p0c
        self._energy0 = math.sqrt(p0c * p0c + self.mass0 * self.mass0)
        self._beta0 = p0c / self._energy0
        self._gamma0 = self._energy0 / self.mass0

    p0c = property(lambda self: self._p0c, lambda self, v: self._set_p0c(float(v)))
    energy0 = property(lambda self: self._energy0)

    @property
    def beta0(self):
        return self._beta0

    @beta0.setter
    def beta0(self, beta0):  # tests/test_particles.py:11-16
        gamma0 = 1.0 / math.sqrt(1.0 - beta0 * beta0)
        self._set_p0c(self.mass0 * beta0 * gamma0)

    @property
    def gamma0(self):
        return self._gamma0

    @gamma0.setter
    def gamma0(self, gamma0):  # tests/test_particles.py:17-19
        self._set_p0c(self.mass0 * math.sqrt(gamma0 * gamma0 - 1.0))

    # -- energy bookkeeping (recalled Pyparticles arithmetic) -------------------------------
    @property
    def delta(self):
        return self._delta

    @delta.setter
    def delta(self, value):
        b0 = self._beta0
        d = torch.as_tensor(value, dtype=torch.float64, device=self.device)
        if d.ndim == 0:
            d = d.expand(len(self)).clone()
        self._delta = d
        db0 = d * b0
        ptaub0 = torch.sqrt(db0 ** 2 + 2 * db0 * b0 + 1) - 1
        opd = 1 + d
        self.rvv = opd / (1 + ptaub0)
        self.rpp = 1 / opd

    def add_to_energy(self, energy):
        b0 = self._beta0
        old = self.rvv
        db0 = self._delta * b0
        ptaub0 = torch.sqrt(db0 ** 2 + 2 * db0 * b0 + 1) - 1
        ptaub0 = ptaub0 + energy / self._energy0
        ptau = ptaub0 / b0
        self._delta = torch.sqrt(ptau ** 2 + 2 * ptau / b0 + 1) - 1
        opd = 1 + self._delta
        self.rvv = opd / (1 + ptaub0)
        self.rpp = 1 / opd
        self.zeta = self.zeta * (self.rvv / old)

    # -- derived longitudinal variables (Pyparticles names; read-only views) -----------------------
    @property
    def ptau(self):
        """(E - E0) / (p0 c)."""
        b0 = self._beta0
        return torch.sqrt(self._delta ** 2 + 2 * self._delta + 1 / (b0 * b0)) - 1 / b0

    @property
    def psigma(self):
        return self.ptau / self._beta0

    @property
    def sigma(self):
        """s - beta0 c t."""
        return self.zeta / self.rvv

    @property
    def tau(self):
        """s / beta0 - c t."""
        return self.zeta / (self.rvv * self._beta0)

    @property
    def energy(self):
        return self._energy0 + self.ptau * self._p0c

    @property
    def pc(self):
        return (1 + self._delta) * self._p0c

    @property
    def mass_ratio(self):
        return self.charge_ratio / self.chi

    # -- container behaviour -----------------------------------------------------------------
    def __len__(self):
        return int(self.x.shape[0])

    def _columns(self):
        for k in FLOAT_COLS:
            yield k, (self._delta if k == "delta" else getattr(self, k))
        for k in INT_COLS:
            yield k, getattr(self, k)

    def _assign(self, k, t):
        if k == "delta":
            self._delta = t
        else:
            setattr(self, k, t)

    def _clone_meta(self):
        new = object.__new__(Particles)
        new.device = self.device
        new.q0, new.mass0 = self.q0, self.mass0
        new._set_p0c(self._p0c)
        new._scalar_input = self._scalar_input
        new._pinned = self._pinned
        new.lost_particles = []
        return new

    def copy(self):
        new = self._clone_meta()
        for k, t in self._columns():
            new._assign(k, t.clone())
        new.lost_particles = [lp.copy() for lp in self.lost_particles]
        return new

    def select(self, mask_or_index):
        new = self._clone_meta()
        for k, t in self._columns():
            new._assign(k, t[mask_or_index].clone())
        return new

    def to(self, device):
        new = self._clone_meta()
        new.device = torch.device(device)
        for k, t in self._columns():
            new._assign(k, t.to(new.device))
        return new

    def remove_lost_particles(self, keep_memory=True):
        """Reference semantics (``tests/test_losses.py:5-17``): drop ``state != 1`` entries
        from every column, order preserved; removed ones go to ``lost_particles``."""
        keep = self.state == 1
        if bool(keep.all()):
            return
        if keep_memory:
            self.lost_particles.append(self.select(~keep))
        for k, t in list(self._columns()):
            self._assign(k, t[keep].clone())

    def compare(self, other, rel_tol=1e-6, abs_tol=1e-15):
        """``p1.compare(p2, abs_tol=...)`` (``tests/test_track.py:45``)."""
        ok = True
        for k in ("x", "px", "y", "py", "zeta", "delta", "s"):
            a = (self._delta if k == "delta" else getattr(self, k)).detach().cpu().numpy()
            b = (other._delta if k == "delta" else getattr(other, k)).detach().cpu().numpy()
            ok = ok and bool(np.all(np.abs(a - b) <= np.maximum(abs_tol, rel_tol * np.maximum(np.abs(a), np.abs(b)))))
        return ok

    def to_numpy(self):
        return {k: t.detach().cpu().numpy().copy() for k, t in self._columns()}

    def __repr__(self):
        return "Particles(n=%d, p0c=%g, device=%s, alive=%d)" % (
            len(self), self._p0c, self.device, int((self.state == 1).sum()))
