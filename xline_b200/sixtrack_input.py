"""SixTrack ``fort.2 / fort.3 / fort.8 / fort.16`` reader and the SixTrack->Line conversion.

The reference's ``Line.from_sixinput`` (``xline/line.py:279-295`` ->
``xline/loader_sixtrack.py:24-162``) consumes a ``sixtracktools.SixInput`` -- a third-party
package that is not vendored, not pinned and not installed here.  :class:`SixInput` below
provides the attributes that loader reads (``iter_struct()``, ``single``, ``align``,
``get_knl(name, count)``, ``bbelements``, ``initialconditions``, ``ition u0 harm tlen pma``)
straight from the shipped input files, following the SixTrack input-file format.  Unit
conventions that cannot be pinned against ``sixtracktools`` (fort.16 scaling, the signs of
the beam-beam separations) are stated where they are applied.  :func:`expand_struct`
restates the element mapping of ``xline/loader_sixtrack.py:24-162``.

Host-side, run once per lattice; nothing here is on the hot path.
"""
import math
import os
from collections import namedtuple

import numpy as np

CLIGHT = 299792458  # xline/loader_sixtrack.py:7

BB4D = namedtuple("BB4D", "charge sigma_x sigma_y beta_r x_bb y_bb d_px d_py")
BB6D = namedtuple(
    "BB6D",
    "phi alpha x_bb_co y_bb_co charge_slices zeta_slices sigma_11 sigma_12 sigma_13 sigma_14 "
    "sigma_22 sigma_23 sigma_24 sigma_33 sigma_34 sigma_44",
)


def _floats(tokens):
    return [float(t.replace("D", "e").replace("d", "e")) for t in tokens]


def _blocks(path):
    """Split a SixTrack input file into ``(keyword_line, [body lines])`` blocks."""
    out = []
    if not os.path.exists(path):
        return out
    with open(path) as fh:
        lines = [ln.rstrip("\n") for ln in fh]
    i = 0
    while i < len(lines):
        head = lines[i].strip()
        i += 1
        if not head or head.startswith("/"):
            continue
        if head.startswith("ENDE"):
            break
        body = []
        while i < len(lines) and lines[i].strip() != "NEXT":
            if lines[i].strip() and not lines[i].lstrip().startswith("/"):
                body.append(lines[i])
            i += 1
        i += 1
        out.append((head, body))
    return out


def constant_charge_slicing_gaussian(n_part_tot, sigmaz, n_slices):
    """Equal-charge longitudinal slices of a Gaussian bunch: slice centroids and charges
    (the construction of ``xline/be_beamfields/slicing.py:5-56``)."""
    from scipy.special import erfinv

    if n_slices < 1:
        raise ValueError("Invalid number of slices")
    if n_slices == 1:
        return np.array([0.0]), np.array([float(n_part_tot)])
    q = np.arange(1, n_slices) / float(n_slices)
    cuts = math.sqrt(2) * sigmaz * erfinv(2 * q - 1.0)
    g = np.exp(-cuts ** 2 / (2 * sigmaz * sigmaz)) * sigmaz / math.sqrt(2 * math.pi) * n_slices
    cent = np.empty(n_slices)
    cent[0] = -g[0]
    cent[1:-1] = -(g[1:] - g[:-1])
    cent[-1] = g[-1]
    return cent, np.full(n_slices, n_part_tot / float(n_slices))


class SixInput:
    """Parsed SixTrack input directory (``fort.2``, ``fort.3``, optional ``fort.8`` /
    ``fort.16``)."""

    def __init__(self, path="."):
        self.path = path
        self.single = {}
        self.blocks = {}
        self.struct = []
        self.align = {}
        self.mult = {}
        self.multblock = {}
        self.bbelements = {}
        self.initialconditions = []
        self.harm = self.alc = self.u0 = self.phag = self.tlen = self.pma = 0.0
        self.ition = 0
        self.beam = None
        self._read_fort2(os.path.join(path, "fort.2"))
        self._read_fort3(os.path.join(path, "fort.3"))
        self._read_fort8(os.path.join(path, "fort.8"))
        self._read_fort16(os.path.join(path, "fort.16"))

    # ---- fort.2: geometry ---------------------------------------------------------------
    def _read_fort2(self, fn):
        if not os.path.exists(fn):
            raise FileNotFoundError(fn)
        for head, body in _blocks(fn):
            if head.startswith("SING"):
                for ln in body:
                    t = ln.split()
                    self.single[t[0]] = [int(t[1])] + _floats(t[2:])
            elif head.startswith("BLOC"):
                cur = None
                for ln in body[1:]:  # first line: mper msym
                    t = ln.split()
                    if not ln[0].isspace():
                        cur = t[0]
                        self.blocks[cur] = []
                        t = t[1:]
                    self.blocks[cur].extend(t)
            elif head.startswith("STRU"):
                for ln in body:
                    self.struct.extend(t for t in ln.split() if t != "GO")

    def iter_struct(self):
        """Single-element names in ring order, blocks expanded."""
        out = []
        for nm in self.struct:
            if nm in self.blocks:
                out.extend(self.blocks[nm])
            else:
                out.append(nm)
        return out

    # ---- fort.3: parameters -------------------------------------------------------------
    def _read_fort3(self, fn):
        for head, body in _blocks(fn):
            key = head[:4]
            if key == "INIT":
                vals = []
                for ln in body:
                    vals.extend(_floats(ln.split()))
                # first line carries 5 control numbers; keep everything, the loader indexes
                # from the end (total energy in MeV is last; xline/loader_sixtrack.py:100)
                self.initialconditions = vals[5:]
            elif key == "SYNC":
                t = _floats(body[0].split())
                self.harm, self.alc, self.u0, self.phag, self.tlen, self.pma = t[:6]
                self.ition = int(t[6])
            elif key == "MULT":
                t = body[0].split()
                rows = [_floats(ln.split()) for ln in body[1:]]
                self.mult[t[0]] = dict(
                    r0=float(t[1]), benda=float(t[2]),
                    bn=[r[0] for r in rows], bnrms=[r[1] for r in rows],
                    an=[r[2] for r in rows], anrms=[r[3] for r in rows],
                )
            elif key == "BEAM":
                self._parse_beam(body)

    def _parse_beam(self, body):
        if not body or body[0].strip() != "EXPERT":
            if body:
                self.beam = _floats(body[0].split())
            return
        hdr = _floats(body[1].split())
        partnum, sigz = hdr[0], hdr[3]
        self.beam = hdr
        i = 2
        while i < len(body):
            t = body[i].split()
            name, nsl = t[0], int(float(t[1]))
            v = _floats(t[2:])
            if nsl == 0:
                # 4D lens: Sxx[mm^2] Syy[mm^2] h-sep[mm] v-sep[mm] strength-ratio
                self.bbelements[name] = BB4D(
                    charge=partnum * v[4], sigma_x=math.sqrt(v[0]) * 1e-3,
                    sigma_y=math.sqrt(v[1]) * 1e-3, beta_r=1.0,
                    x_bb=v[2] * 1e-3, y_bb=v[3] * 1e-3, d_px=0.0, d_py=0.0)
                i += 1
            else:
                # 6D lens: xang xplane h-sep v-sep / Sxx Sxxp Sxpxp Syy Syyp /
                #          Sypyp Sxy Sxyp Sxpy Sxpyp strength-ratio  (mm, mrad units)
                a = _floats(body[i + 1].split())
                b = _floats(body[i + 2].split())
                zc, npart = constant_charge_slicing_gaussian(partnum * b[5], sigz, nsl)
                self.bbelements[name] = BB6D(
                    phi=v[0], alpha=v[1], x_bb_co=v[2] * 1e-3, y_bb_co=v[3] * 1e-3,
                    charge_slices=list(npart), zeta_slices=list(zc),
                    sigma_11=a[0] * 1e-6, sigma_12=a[1] * 1e-6, sigma_22=a[2] * 1e-6,
                    sigma_33=a[3] * 1e-6, sigma_34=a[4] * 1e-6, sigma_44=b[0] * 1e-6,
                    sigma_13=b[1] * 1e-6, sigma_14=b[2] * 1e-6, sigma_23=b[3] * 1e-6,
                    sigma_24=b[4] * 1e-6)
                i += 3

    # ---- fort.8: misalignments, one line per occurrence -------------------------------
    def _read_fort8(self, fn):
        if not os.path.exists(fn):
            return
        with open(fn) as fh:
            for ln in fh:
                t = ln.split()
                if len(t) >= 4:
                    self.align.setdefault(t[0], []).append(tuple(_floats(t[1:4])))

    # ---- fort.16: multipole errors, name line + 40 numbers (bn1 an1 bn2 an2 ...) -----
    def _read_fort16(self, fn):
        if not os.path.exists(fn) or os.path.getsize(fn) == 0:
            return
        with open(fn) as fh:
            toks = fh.read().split()
        i = 0
        while i < len(toks):
            name = toks[i]
            vals = _floats(toks[i + 1: i + 41])
            self.multblock.setdefault(name, []).append((vals[0::2], vals[1::2]))
            i += 41

    def synthesize_fort16(self, seed=0, kick_at_r0=1e-8, decade_per_orders=5.0):
        """Deterministic stand-in for a missing ``fort.16`` (the LHC one is a stripped
        large blob, ``.MISSING_LARGE_BLOBS:1``): Gaussian random multipole errors for every
        occurrence of every type-11 element that has a MULT block, scaled so that each
        order n contributes an r.m.s. kick of ``kick_at_r0 * 10**(-(n-1)/decade_per_orders)``
        rad at the reference radius (field errors fall off with multipole order)."""
        rng = np.random.default_rng(seed)
        counts = {}
        for nm in self.iter_struct():
            if nm in self.mult and self.single[nm][0] == 11:
                counts[nm] = counts.get(nm, 0) + 1
        for nm in sorted(counts):
            m = self.mult[nm]
            nord = len(m["bn"])
            d0 = abs(m["benda"]) if m["benda"] != 0 else 1.0
            scale = kick_at_r0 / (d0 * 1e-3) * 10.0 ** (-np.arange(nord) / decade_per_orders)
            self.multblock[nm] = [
                (list(rng.normal(0, 1, nord) * scale), list(rng.normal(0, 1, nord) * scale))
                for _ in range(counts[nm])
            ]

    def get_knl(self, name, count=0):
        """Integrated normal/skew strengths of a type-11 multipole occurrence: the fort.16
        values (relative errors) scaled by the fort.3 MULT block (r.m.s. columns, reference
        radius r0 [mm], bending strength) through ``bn_rel`` of
        ``xline/loader_sixtrack.py:15-21``; normal sign -1, skew sign +1 as for the single
        multipoles at ``xline/loader_sixtrack.py:75-83``."""
        if name not in self.mult:
            return [0.0], [0.0]
        m = self.mult[name]
        nord = len(m["bn"])
        blocks = self.multblock.get(name, [])
        if count < len(blocks):
            bn16, an16 = blocks[count]
        else:
            bn16, an16 = [0.0] * nord, [0.0] * nord
        r0, d0 = m["r0"], m["benda"]
        knl = bn_rel(bn16[:nord], m["bnrms"], r0, d0, -1)
        ksl = bn_rel(an16[:nord], m["anrms"], r0, d0, +1)
        return knl, ksl


def bn_mad(bn, n, sign):
    """xline/loader_sixtrack.py:11-12."""
    return sign * bn * math.factorial(n - 1)


def bn_rel(bn16, bn3, r0, d0, sign):
    """xline/loader_sixtrack.py:15-21."""
    out = []
    for nn, (a, b) in enumerate(zip(bn16, bn3)):
        n = nn + 1
        sixval = d0 * a * b * r0 ** (1 - n) * 10 ** (3 * n - 6)
        out.append(bn_mad(sixval, n, sign))
    return out


def expand_struct(six, classes):
    """SixTrack structure -> ``[(name, type_name, element)]``, ``rest``, ``iconv``
    (restates ``xline/loader_sixtrack.py:24-162``; ``classes`` maps names to element
    classes, the reference's ``convert=`` namespace)."""
    if not isinstance(classes, dict):
        classes = {k: getattr(classes, k) for k in dir(classes) if not k.startswith("_")}
    Drift, Multipole, Cavity = classes["Drift"], classes["Multipole"], classes["Cavity"]
    XYShift, SRotation = classes["XYShift"], classes["SRotation"]
    RFMultipole = classes["RFMultipole"]
    out, rest, iconv = [], [], []
    occurrence = {}
    icount = 0
    names = six.iter_struct()
    if "CAV" in names:  # loader_sixtrack.py:42-48
        six.single["CAV"] = [12 * six.ition, six.u0, six.harm, 0]
    for nm in names:
        occ = occurrence.get(nm, 0)
        etype, d1, d2, d3 = six.single[nm][:4]
        elem, exclude = None, False
        shift = tilt = None
        if nm in six.align and occ < len(six.align[nm]):  # loader_sixtrack.py:57-71
            dx, dy, tl = six.align[nm][occ]
            tl = tl * 180e-3 / math.pi
            dx, dy = dx * 1e-3, dy * 1e-3
            if abs(dx) + abs(dy) > 0:
                shift = (dx, dy)
                out.append((nm + "_preshift", "XYShift", XYShift(dx=dx, dy=dy)))
                icount += 1
            if abs(tl) > 0:
                tilt = tl
                out.append((nm + "_pretilt", "SRotation", SRotation(angle=tl)))
                icount += 1
        if etype in (0, 25):
            elem = Drift(length=d3)
            exclude = d3 > 0
        elif abs(etype) in (1, 2, 3, 4, 5, 7, 8, 9, 10):
            nn = abs(etype)
            sign = -etype / nn
            val = bn_mad(d1, nn, sign)
            knl, ksl = [0] * (nn - 1) + [val], [0] * nn
            if sign == 1:
                knl, ksl = ksl, knl
            elem = Multipole(knl=knl, ksl=ksl, hxl=0, hyl=0, length=0)
        elif etype == 11:
            knl, ksl = six.get_knl(nm, occ)
            knl, ksl = list(knl), list(ksl)
            hxl = hyl = length = 0
            if d3 == -1:
                hxl, length = -d1, d2
                knl[0] = hxl
            elif d3 == -2:
                hyl, length = -d1, d2
                ksl[0] = hyl
            elem = Multipole(knl=knl, ksl=ksl, hxl=hxl, hyl=hyl, length=length)
        elif etype == 12:
            e0 = six.initialconditions[-1]
            p0c = math.sqrt(e0 ** 2 - six.pma ** 2)
            beta0 = p0c / e0
            elem = Cavity(voltage=d1 * 1e6, frequency=d2 * CLIGHT * beta0 / six.tlen, lag=180 - d3)
        elif etype == 20:
            bb = six.bbelements[nm]
            if hasattr(bb, "sigma_x"):
                elem = classes["BeamBeam4D"](**bb._asdict())
            elif hasattr(bb, "phi"):
                elem = classes["BeamBeam6D"](**bb._asdict())
            else:
                raise ValueError("What?!")  # loader_sixtrack.py:117
        elif etype in (23, -23):
            p0c_eV = six.initialconditions[12] * 1e6
            if etype == 23:
                elem = RFMultipole(frequency=d2 * 1e6, knl=[d1 * 1e6 / p0c_eV], pn=[90.0])
            else:
                elem = RFMultipole(frequency=d2 * 1e6, ksl=[-d1 * 1e6 / p0c_eV], ps=[90.0])
        else:
            rest.append([nm] + list(six.single[nm]))
        if elem is not None:
            out.append((nm, type(elem).__name__, elem))
        if tilt is not None:
            out.append((nm + "_posttilt", "SRotation", SRotation(angle=-tilt)))
            icount += 1
        if shift is not None:
            out.append((nm + "_postshift", "XYShift", XYShift(dx=-shift[0], dy=-shift[1])))
            icount += 1
        if elem is not None:
            if not exclude:
                iconv.append(icount)
            icount += 1
        occurrence[nm] = occ + 1
    return out, rest, iconv
