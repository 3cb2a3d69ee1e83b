"""The benchmark configurations of BASELINE.json rebuilt from this package's serialised
lattices (``xline_b200/lattices/*.json.gz``, converted from the reference's SixTrack
example inputs by ``scripts/import_reference_lattices.py``) plus the synthetic beams of
SURVEY.md §8(d).  Host-side setup; used by ``bench.py``, ``smoke()`` and the tests.
"""
import gzip
import json
import os

import numpy as np

from . import elements as E
from .line import Line

_HERE = os.path.dirname(os.path.abspath(__file__))
SEED0 = 20261018


def load_lattice(name):
    """-> (Line, meta dict)."""
    fn = os.path.join(_HERE, "lattices", name + ".json.gz")
    with gzip.open(fn, "rb") as fh:
        d = json.loads(fh.read().decode())
    return Line.from_dict(d), d.get("meta", {})


def p0c_of(meta):
    e0, m0 = meta["energy0_eV"], meta["mass0_eV"]
    return float(np.sqrt(e0 * e0 - m0 * m0)), float(m0)


def add_lhc_apertures(line, rect=(0.022, 0.018), ellipse=(0.022, 0.018)):
    """C2 of SURVEY.md §8(d): a ``LimitRect`` after every high-order (SixTrack type-11)
    multipole and a ``LimitEllipse`` after every quadrupole (order-1 multipole)."""
    els, names = [], []
    for el, nm in zip(line.elements, line.element_names):
        els.append(el)
        names.append(nm)
        if isinstance(el, E.Multipole):
            if el.order >= 4:
                els.append(E.LimitRect(min_x=-rect[0], max_x=rect[0], min_y=-rect[1], max_y=rect[1]))
                names.append(nm + "_aper")
            elif el.order == 1:
                els.append(E.LimitEllipse(a=ellipse[0], b=ellipse[1]))
                names.append(nm + "_aper")
    return Line(els, names)


def gaussian_beam(n, config, rank=0, sx=1e-4, spx=1e-6, sz=0.077, sd=1.1e-4, amp_max=4.0,
                  first_id=0):
    """Synthetic beam of SURVEY.md §8(d): Gaussian core scaled per particle by an
    amplitude factor A ~ U(0, amp_max) (A = 1 when amp_max is None)."""
    rng = np.random.default_rng(SEED0 + 1000 * config + rank)
    a = rng.uniform(0.0, amp_max, n) if amp_max else np.ones(n)
    return dict(
        x=rng.normal(0, sx, n) * a, px=rng.normal(0, spx, n) * a,
        y=rng.normal(0, sx, n) * a, py=rng.normal(0, spx, n) * a,
        zeta=rng.normal(0, sz, n), delta=rng.normal(0, sd, n),
        particle_id=np.arange(first_id, first_id + n),
    )


def config_fodo(n=10_000, rank=0):
    """C1: examples/fodo FODO cell, 10k particles x 100 turns."""
    line, meta = load_lattice("fodo")
    p0c, m0 = p0c_of(meta)
    rng = np.random.default_rng(SEED0 + 1000 * 1 + rank)
    cols = dict(x=rng.normal(0, 1e-3, n), px=rng.normal(0, 1e-4, n), y=rng.normal(0, 1e-3, n),
                py=rng.normal(0, 1e-4, n), zeta=rng.normal(0, 0.05, n), delta=rng.normal(0, 3e-4, n))
    return line, cols, p0c, m0


def config_lhc(n=1_000_000, rank=0, apertures=True, first_id=0):
    """C2: LHC lattice (examples/lhc), synthetic apertures, Gaussian beam (core sigma
    1e-4 m / 1e-6 rad at IP3) with per-particle amplitude scale A ~ U(0, 4): about 90 % of
    the particles survive the first few hundred turns, the rest hit the apertures."""
    line, meta = load_lattice("lhc")
    if apertures:
        line = add_lhc_apertures(line)
    p0c, m0 = p0c_of(meta)
    return line, gaussian_beam(n, 2, rank, first_id=first_id), p0c, m0


def config_lhc_beambeam(n=10_000_000, rank=0, first_id=0):
    """C3: LHC with BeamBeam4D/6D lenses (examples/beambeam)."""
    line, meta = load_lattice("lhc_beambeam")
    p0c, m0 = p0c_of(meta)
    return line, gaussian_beam(n, 3, rank, sx=2e-4, spx=2e-6, sz=0.075, amp_max=6.0,
                               first_id=first_id), p0c, m0


# ---------------------------------------------------------------------------------------
# Synthetic stand-ins for the MAD-X based configurations (C4 PETRA IV, C5 PSB): the
# reference builds those lattices with cpymad + MAD-X `makethin`, neither of which exists
# here.  The generators below reproduce the element *mix* and sizes the
# survey extracted from examples/petra4/h7ba_n8.seq and tests/psb/*, on simple stable
# optics, so that parity and throughput of those element types can be measured.
# ---------------------------------------------------------------------------------------
ELECTRON_MASS_EV = 0.51099895e6


def config_psb_like(n=1_000_000, rank=0, n_cells=16, sc_per_cell=8, monitor_stores=0, monitor_ids=0):
    """C5 stand-in: a 157 m proton ring at p0c = 0.571 GeV (tests/psb/psb_fb_lhc.madx:14) made
    of 16 FODO cells with curved thin dipoles, DipoleEdge pairs, 128 SCQGaussProfile kicks per
    turn (reference: 120, tests/test_madx_import.py:19; 1e11 protons, bunchlength_rms = 1 m,
    :30-33), the PSB aperture mix (circle / ellipse / rect-ellipse / rectangle) and a h = 1
    cavity; optionally one BeamMonitor."""
    from .particles import PROTON_MASS_EV

    p0c, m0 = 0.571e9, PROTON_MASS_EV
    circ = 157.08
    lc = circ / n_cells
    kq = 0.24  # thin-quad strength [1/m]: ~72 degrees per cell
    bend = 2 * np.pi / (2 * n_cells)
    seg = lc / 2 / (sc_per_cell // 2)  # drift between space-charge kicks
    els, names = [], []

    def add(el, nm):
        els.append(el)
        names.append(nm)

    beta_max, beta_min = 16.0, 5.0
    eps = 2.5e-6 / (0.608 * 1.17)  # normalised 2.5 um at beta*gamma = 0.712
    for c in range(n_cells):
        for half, (k1l, bx, by) in enumerate(((kq, beta_max, beta_min), (-kq, beta_min, beta_max))):
            tag = "c%d.%s" % (c, "f" if half == 0 else "d")
            add(E.Multipole(knl=[0.0, k1l], ksl=[0.0, 0.0]), "q" + tag)
            ap_kind = (c + half) % 4
            if ap_kind == 0:
                add(E.LimitEllipse(a=0.05, b=0.05), "ap" + tag)          # circle
            elif ap_kind == 1:
                add(E.LimitEllipse(a=0.06, b=0.035), "ap" + tag)
            elif ap_kind == 2:
                add(E.LimitRectEllipse(max_x=0.05, max_y=0.03, a=0.06, b=0.04), "ap" + tag)
            else:
                add(E.LimitRect(min_x=-0.055, max_x=0.055, min_y=-0.032, max_y=0.032), "ap" + tag)
            for k in range(sc_per_cell // 2):
                add(E.Drift(length=seg / 2), "d%s.%da" % (tag, k))
                f = (k + 0.5) / (sc_per_cell // 2)
                sx = float(np.sqrt(eps * (bx + (by - bx) * f)))
                sy = float(np.sqrt(eps * (by + (bx - by) * f)))
                add(E.SCQGaussProfile(number_of_particles=1e11, bunchlength_rms=1.0, sigma_x=sx,
                                      sigma_y=sy, length=seg, x_co=0.0, y_co=0.0), "sc%s.%d" % (tag, k))
                if k == 0:
                    add(E.DipoleEdge(h=bend / 1.6, e1=bend / 2, hgap=0.03, fint=0.5), "e1" + tag)
                    add(E.Multipole(knl=[bend], ksl=[0.0], hxl=bend, hyl=0.0, length=1.6), "b" + tag)
                    add(E.DipoleEdge(h=bend / 1.6, e1=bend / 2, hgap=0.03, fint=0.5), "e2" + tag)
                add(E.Drift(length=seg / 2), "d%s.%db" % (tag, k))
        if c == 0 and monitor_stores > 0:
            add(E.BeamMonitor(num_stores=monitor_stores, start=0, skip=1, min_particle_id=0,
                              max_particle_id=max(monitor_ids - 1, 0)), "monitor")
    e0 = np.sqrt(p0c ** 2 + m0 ** 2)
    frev = (p0c / e0) * 299792458.0 / circ
    add(E.Cavity(voltage=8e3, frequency=frev, lag=0.0), "cav")
    rng = np.random.default_rng(SEED0 + 1000 * 5 + rank)
    cols = dict(
        x=rng.normal(0, 4e-3, n), px=rng.normal(0, 4e-4, n), y=rng.normal(0, 3e-3, n),
        py=rng.normal(0, 3e-4, n), zeta=rng.normal(0, 1.0, n), delta=rng.normal(0, 1e-3, n),
    )
    return Line(els, names), cols, p0c, m0


def config_petra_like(n=1_000_000, rank=0, n_cells=120, grid=None):
    """C4 stand-in: a 6 GeV electron ring (examples/petra4/track_p1.py:21) of 120 cells whose
    quadrupoles and bends are 4-slice thin lenses (track_p1.py:26-30) separated by exact
    drifts, with sextupoles, two 500 MHz cavities and two RFMultipoles (h7ba_n8.seq:259,918,
    923 for the RF) -- about 5.6k elements per turn.  The beam is a dynamic-aperture scan: a
    2-D grid of (x, y) amplitudes, interleaved across ranks when sharded."""
    p0c, m0 = 6e9, ELECTRON_MASS_EV
    lc = 2304.0 / n_cells
    kq = 4 * np.sin(np.radians(40.0)) / lc  # 80 degrees per cell
    bend = 2 * np.pi / (2 * n_cells)
    els, names = [], []

    def add(el, nm):
        els.append(el)
        names.append(nm)

    dq = 0.05  # drift between slices
    dl = (lc / 2 - 4 * dq * 2) / 2
    for c in range(n_cells):
        for half, sgn in enumerate((1.0, -1.0)):
            tag = "%d.%d" % (c, half)
            for sl in range(4):
                add(E.Multipole(knl=[0.0, sgn * kq / 4], ksl=[0.0, 0.0]), "q%s.%d" % (tag, sl))
                add(E.DriftExact(length=dq), "dq%s.%d" % (tag, sl))
            add(E.Multipole(knl=[0.0, 0.0, sgn * 1.5], ksl=[0.0, 0.0, 0.0]), "s" + tag)
            add(E.DriftExact(length=dl), "d1" + tag)
            for sl in range(4):
                add(E.Multipole(knl=[bend / 4], ksl=[0.0], hxl=bend / 4, hyl=0.0, length=0.4),
                    "b%s.%d" % (tag, sl))
                add(E.DriftExact(length=dq), "db%s.%d" % (tag, sl))
            add(E.DriftExact(length=dl), "d2" + tag)
        if c in (0, n_cells // 2):
            add(E.Cavity(voltage=4e6, frequency=499.6e6, lag=180.0), "cav%d" % c)
            add(E.RFMultipole(voltage=0.0, frequency=499.6e6, lag=0.0, knl=[0.0, 1e-4], ksl=[0.0, 0.0],
                              pn=[0.0, 90.0], ps=[0.0, 0.0]), "rfm%d" % c)
            add(E.LimitEllipse(a=0.01, b=0.005), "ap%d" % c)
    if grid is None:
        side = int(np.ceil(np.sqrt(n)))
        gx, gy = np.meshgrid(np.linspace(0, 4e-3, side), np.linspace(0, 2e-3, side))
        x0, y0 = gx.ravel()[:n], gy.ravel()[:n]
    else:
        x0, y0 = grid
    cols = dict(x=x0.copy(), px=np.zeros(n), y=y0.copy(), py=np.zeros(n), zeta=np.zeros(n),
                delta=np.zeros(n))
    return Line(els, names), cols, p0c, m0


def install_spacecharge(line, n_kicks, p0c, mass0, number_of_particles, bunchlength_rms, neps_x, neps_y,
                        delta_rms, kind="bunched", circumference=None):
    """Cut ``n_kicks`` equidistant space-charge kicks into the drifts of ``line`` and size
    them from the line's own lattice functions -- the recipe the reference's PSB test
    prepares (tests/test_madx_import.py:19-36: ``n_SCkicks = 120``, 1e11 protons,
    ``bunchlength_rms = 1``, normalised emittances 1.5 um, ``delta_rms = 1e-3``), with
    ``optics.twiss`` in place of the MAD-X twiss table:
    ``sigma_x = sqrt(betx eps_x / (beta gamma) + (dx delta_rms)^2)`` at every kick, kick length
    = circumference / n_kicks.  Returns the new Line."""
    from . import optics

    circ = line.get_length()
    seg = circ / n_kicks
    targets = [(k + 0.5) * seg for k in range(n_kicks)]
    els, names = [], []
    s, t = 0.0, 0
    sc_ids = []
    for el, nm in zip(line.elements, line.element_names):
        if isinstance(el, (E.Drift, E.DriftExact)) and el.length > 0:
            cls, rest, part = type(el), el.length, 0
            while t < n_kicks and targets[t] <= s + rest:
                head = targets[t] - s
                if head > 0:
                    els.append(cls(length=head))
                    names.append("%s..%d" % (nm, part))
                    part += 1
                sc_ids.append(len(els))
                els.append(None)
                names.append("sc_%d" % t)
                s, rest, t = targets[t], rest - head, t + 1
            if rest > 0 or part == 0:
                els.append(cls(length=rest))
                names.append(nm if part == 0 else "%s..%d" % (nm, part))
            s += rest
        else:
            els.append(el)
            names.append(nm)
    assert t == n_kicks, "could not place every space-charge kick inside a drift"
    bare = Line([e if e is not None else E.Drift(length=0.0) for e in els], names)
    tw = optics.twiss(bare)
    bg = p0c / mass0
    for i in sc_ids:
        sx = float(np.sqrt(tw["betx"][i] * neps_x / bg + (tw["dx"][i] * delta_rms) ** 2))
        sy = float(np.sqrt(tw["bety"][i] * neps_y / bg + (tw["dy"][i] * delta_rms) ** 2))
        if kind == "bunched":
            els[i] = E.SCQGaussProfile(number_of_particles=number_of_particles, bunchlength_rms=bunchlength_rms,
                                       sigma_x=sx, sigma_y=sy, length=seg, x_co=0.0, y_co=0.0)
        else:
            els[i] = E.SCCoasting(number_of_particles=number_of_particles,
                                  circumference=circumference or circ, sigma_x=sx, sigma_y=sy,
                                  length=seg, x_co=0.0, y_co=0.0)
    return Line(els, names)


def matched_gaussian(line, n, rng, p0c, mass0, neps_x, neps_y, delta_rms, sigma_zeta):
    """Gaussian beam matched to the lattice functions at the start of ``line``."""
    from . import optics

    tw = optics.twiss(line)
    bg = p0c / mass0
    delta = rng.normal(0, delta_rms, n)
    cols = dict(zeta=rng.normal(0, sigma_zeta, n), delta=delta)
    for u, pu, eps in (("x", "px", neps_x / bg), ("y", "py", neps_y / bg)):
        beta, alpha = tw["bet" + u][0], tw["alf" + u][0]
        g1, g2 = rng.normal(0, 1, n), rng.normal(0, 1, n)
        cols[u] = np.sqrt(eps * beta) * g1 + tw["d" + u][0] * delta
        cols[pu] = np.sqrt(eps / beta) * (g2 - alpha * g1) + tw["d" + pu][0] * delta
    return cols


def eta_sign(line, p0c, m0):
    """Sign of the slip factor ``alfa_c - 1 / gamma0^2``."""
    from . import optics

    e0 = np.sqrt(p0c ** 2 + m0 ** 2)
    return np.sign(optics.twiss(line)["alfa_c"] - (m0 / e0) ** 2)


def config_psb(n=1_000_000, rank=0, n_sc=120, monitor_stores=0, monitor_ids=0, number_of_particles=1e11,
               bunchlength_rms=1.0, rf_voltage=None, monitor_skip=1):
    """C5: the PS Booster ring 1 of tests/psb/psb_fb_lhc.madx (flat-bottom optics,
    PC = 0.571 GeV/c, QH = 4.22, QV = 4.45; 380 thin elements + 264 apertures from
    psb_aperture.dbx), read by ``xline_b200.madx_input``, with 120 ``SCQGaussProfile`` kicks
    (tests/test_madx_import.py:19-33) sized from ``optics.twiss``, an h = 1 RF voltage on the
    first ``ACWFB`` cavity (the shipped strength file leaves it at 0; by default the voltage
    that matches ``bunchlength_rms`` to ``delta_rms = 1e-3`` in the linear bucket, so that the
    tracked bunch keeps the length the frozen space-charge profile assumes) and optionally one
    BeamMonitor at the start of the ring.  Beam: Gaussian matched to the bare optics."""
    line, meta = load_lattice("psb")
    p0c, m0 = p0c_of(meta)
    neps, delta_rms = 1.5e-6, 1e-3
    line = install_spacecharge(line, n_sc, p0c, m0, number_of_particles, bunchlength_rms, neps, neps, delta_rms)
    e0 = np.sqrt(p0c ** 2 + m0 ** 2)
    frev = (p0c / e0) * 299792458.0 / line.get_length()
    if rf_voltage is None:
        from . import optics

        beta0, gamma0 = p0c / e0, e0 / m0
        eta = optics.twiss(line)["alfa_c"] - 1.0 / gamma0 ** 2
        radius = line.get_length() / (2 * np.pi)
        qs = abs(eta) * radius * delta_rms / bunchlength_rms       # sigma_z = |eta| R sigma_delta / Qs
        rf_voltage = 2 * np.pi * beta0 ** 2 * e0 * qs ** 2 / abs(eta)   # Qs^2 = h |eta| V / (2 pi beta^2 E), h = 1
    cav = [e for e in line.elements if isinstance(e, E.Cavity)][0]
    # below transition (eta < 0): stable phase 0, xline lag in degrees (xline/elements.py:241)
    cav.voltage, cav.frequency, cav.lag = float(rf_voltage), float(frev), 0.0 if eta_sign(line, p0c, m0) < 0 else 180.0
    if monitor_stores > 0:
        line.insert_element(0, E.BeamMonitor(num_stores=monitor_stores, start=0, skip=monitor_skip,
                                             min_particle_id=0, max_particle_id=max(monitor_ids - 1, 0)), "monitor")
    rng = np.random.default_rng(SEED0 + 1000 * 5 + rank)
    cols = matched_gaussian(line, n, rng, p0c, m0, neps, neps, delta_rms, bunchlength_rms)
    return line, cols, p0c, m0


def config_petra4(n=1_000_000, rank=0, x_max=1.5e-3, y_max=0.8e-3):
    """C4: the PETRA IV lattice of examples/petra4/h7ba_n8.seq (4 886 placements; 2 rfcavity at
    500 MHz), read by ``xline_b200.madx_input`` and made thin with 4 TEAPOT slices per
    quadrupole / bend (examples/petra4/track_p1.py:26-30), exact drifts: 31 025 elements.
    Beam: a dynamic-aperture scan, a 2-D grid of (x, y) start amplitudes."""
    line, meta = load_lattice("petra4")
    p0c, m0 = p0c_of(meta)
    side = int(np.ceil(np.sqrt(n)))
    gx, gy = np.meshgrid(np.linspace(0, x_max, side), np.linspace(0, y_max, side))
    cols = dict(x=gx.ravel()[:n].copy(), px=np.zeros(n), y=gy.ravel()[:n].copy(), py=np.zeros(n),
                zeta=np.zeros(n), delta=np.zeros(n))
    return line, cols, p0c, m0
