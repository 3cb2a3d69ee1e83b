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
