"""MAD-X sequence import without cpymad / MAD-X.

The reference's ``Line.from_madx_sequence`` (``xline/line.py:297-324`` ->
``xline/loader_mad.py:6-249``) iterates a *cpymad* sequence that MAD-X has already expanded
and made thin.  Neither cpymad nor MAD-X exists here, so this module provides the two
missing pieces for the subset of the MAD-X language the shipped lattice files use
(``examples/petra4/h7ba_n8.seq``: scalar variables with deferred expressions, element
definitions with class inheritance, one flat ``sequence`` with ``at=`` positions):

* :class:`MadxFile` -- parser + expression evaluator, giving the thick sequence;
* :func:`makethin` -- TEAPOT thin slicing of quadrupoles / bends (``n`` slices: end drifts
  ``L/(2(n+1))``, inner drifts ``L n/(n^2-1)``), single centre kicks for sextupoles and
  octupoles, ``dipedge`` elements at bend faces -- the element types MAD-X ``makethin``
  hands to the reference's loader;
* :func:`iter_from_madx_sequence` -- the element mapping of ``xline/loader_mad.py:23-249``
  restated for this package's classes.  The objects produced by :func:`makethin` quack like
  cpymad elements (``name``, ``base_type.name``, attributes, ``element_positions()``), so the
  reference's own loader can be run on them: that is how the mapping is pinned
  (``tests/test_madx_import.py``).

Host-side setup, not on the hot path.  Unpinned: agreement with MAD-X's own ``makethin``
output (no MAD-X binary to compare with).
"""
import math
import re


class _BaseType:
    def __init__(self, name):
        self.name = name


class MadElement:
    """A (thick or thin) element instance; attribute access like a cpymad element."""

    def __init__(self, name, base_type, attrs, position=0.0):
        self.name = name
        self.base_type = _BaseType(base_type)
        self.position = position  # entry position [m]
        for k, v in attrs.items():
            setattr(self, k, v)
        if not hasattr(self, "l"):
            self.l = 0.0

    def __repr__(self):
        return "MadElement(%s: %s @ %.6f)" % (self.name, self.base_type.name, self.position)


class MadSequence:
    def __init__(self, name, length, elements):
        self.name = name
        self.length = length
        self.elements = elements

    def element_positions(self):
        return [e.position for e in self.elements]


_FUNCS = {k: getattr(math, k) for k in
          ("sin", "cos", "tan", "asin", "acos", "atan", "sqrt", "exp", "log", "sinh", "cosh", "tanh")}
_FUNCS.update({"abs": abs, "pi": math.pi, "twopi": 2 * math.pi, "clight": 299792458.0, "e": math.e})
_BASE_TYPES = {
    "quadrupole", "sbend", "rbend", "sextupole", "octupole", "marker", "monitor", "hmonitor",
    "vmonitor", "instrument", "drift", "hkicker", "vkicker", "kicker", "tkicker", "rfcavity",
    "multipole", "dipedge", "collimator", "rcollimator", "solenoid", "placeholder", "sequence",
}


class MadxFile:
    """Parser for the MAD-X subset described in the module docstring."""

    def __init__(self, path=None, text=None):
        if text is None:
            with open(path) as fh:
                text = fh.read()
        self.vars = {}        # name -> expression string (deferred) or float
        self.elements = {}    # name -> (parent, {attr: expr})
        self.sequences = {}   # name -> (length expr, [(elem name, at expr)])
        self._cache = {}
        self._parse(text)

    # ---- parsing -----------------------------------------------------------------------
    @staticmethod
    def _statements(text):
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        lines = []
        for ln in text.splitlines():
            ln = ln.split("!")[0].split("//")[0]
            lines.append(ln)
        for st in " ".join(lines).split(";"):
            st = st.strip()
            if st:
                yield st.lower()

    def _parse(self, text):
        cur_seq = None
        for st in self._statements(text):
            if st == "endsequence":
                cur_seq = None
                continue
            m = re.match(r"^([\w.]+)\s*:\s*([\w.]+)\s*(?:,(.*))?$", st)
            if m and "=" not in m.group(1):
                name, parent, rest = m.group(1), m.group(2), m.group(3) or ""
                attrs = self._attrs(rest)
                if parent == "sequence":
                    cur_seq = name
                    self.sequences[name] = (attrs.get("l", "0"), [])
                else:
                    self.elements[name] = (parent, attrs)
                continue
            if cur_seq is not None:
                m = re.match(r"^([\w.]+)\s*,\s*at\s*:?=\s*(.+)$", st)
                if m:
                    self.sequences[cur_seq][1].append((m.group(1), m.group(2).strip()))
                    continue
            m = re.match(r"^(?:const\s+|real\s+)*([\w.]+)\s*:?=\s*(.+)$", st)
            if m:
                self.vars[m.group(1)] = m.group(2).strip()
                continue
            # commands (beam, use, call, ...) are ignored

    @staticmethod
    def _attrs(rest):
        out = {}
        depth, cur, parts = 0, "", []
        for ch in rest:
            if ch in "({":
                depth += 1
            elif ch in ")}":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        if cur.strip():
            parts.append(cur)
        for part in parts:
            m = re.match(r"^\s*([\w.]+)\s*:?=\s*(.+?)\s*$", part)
            if m:
                out[m.group(1)] = m.group(2)
        return out

    # ---- evaluation --------------------------------------------------------------------
    def value(self, expr, _stack=()):
        """Evaluate an expression (deferred variables resolved recursively)."""
        if isinstance(expr, (int, float)):
            return float(expr)
        expr = expr.strip()
        if expr in self._cache:
            return self._cache[expr]
        names = set(re.findall(r"[a-z_][\w.]*", expr))
        env = dict(_FUNCS)
        pyexpr = expr.replace("^", "**")
        for nm in names:
            if nm in _FUNCS:
                continue
            if nm in _stack:
                raise ValueError("circular definition of %s" % nm)
            if re.fullmatch(r"e[+-]?\d*", nm) or re.fullmatch(r"\d", nm[:1]):
                continue
            val = self.value(self.vars[nm], _stack + (nm,)) if nm in self.vars else 0.0
            safe = re.sub(r"\W", "_", nm)
            if safe != nm:
                pyexpr = re.sub(r"(?<![\w.])" + re.escape(nm) + r"(?![\w.])", safe, pyexpr)
            env[safe] = val
        if pyexpr.startswith("{") and pyexpr.endswith("}"):
            res = [self.value(t, _stack) for t in pyexpr[1:-1].split(",") if t.strip()]
        else:
            res = float(eval(pyexpr, {"__builtins__": {}}, env))  # noqa: S307 (lattice file arithmetic)
        self._cache[expr] = res
        return res

    def element_attrs(self, name):
        """(base type, merged attribute values) following the class inheritance chain."""
        chain = []
        cur = name
        while cur in self.elements:
            parent, attrs = self.elements[cur]
            chain.append(attrs)
            cur = parent
        if cur not in _BASE_TYPES:
            raise ValueError('MAD element "%s" not recognized' % cur)
        merged = {}
        for attrs in reversed(chain):
            merged.update(attrs)
        return cur, {k: self.value(v) for k, v in merged.items()}

    def sequence(self, name):
        """The thick sequence: elements at their ENTRY positions, centre-referred ``at``."""
        length_expr, placements = self.sequences[name]
        out = []
        for ename, at in placements:
            base, attrs = self.element_attrs(ename)
            centre = self.value(at)
            length = float(attrs.get("l", 0.0))
            out.append(MadElement(ename, base, attrs, centre - 0.5 * length))
        return MadSequence(name, self.value(length_expr), out)


def _teapot(length, n):
    """Kick positions (from the entry) of an n-slice TEAPOT thin-lens model."""
    if n == 1:
        return [0.5 * length]
    end = length / (2.0 * (1 + n))
    inner = length * n / (n * n - 1.0)
    return [end + i * inner for i in range(n)]


def makethin(seq, slices=None, default_slices=1):
    """Thin version of a thick :class:`MadSequence` (what MAD-X ``makethin`` would hand to the
    loader): ``slices`` maps a base type to its number of TEAPOT slices
    (``examples/petra4/track_p1.py:26-30`` uses 4 for ``sbend`` and ``quadrupole``)."""
    slices = dict(slices or {})
    out = []
    for el in seq.elements:
        base, L = el.base_type.name, float(el.l)
        n = int(slices.get(base, default_slices))
        tilt = float(getattr(el, "tilt", 0.0))
        if base in ("quadrupole", "sextupole", "octupole") and L > 0:
            order = {"quadrupole": 1, "sextupole": 2, "octupole": 3}[base]
            strength = float(getattr(el, "k%d" % order, 0.0))
            skew = float(getattr(el, "k%ds" % order, 0.0))
            for i, s in enumerate(_teapot(L, n)):
                knl = [0.0] * order + [strength * L / n]
                ksl = [0.0] * order + [skew * L / n]
                out.append(MadElement("%s..%d" % (el.name, i + 1) if n > 1 else el.name, "multipole",
                                      dict(knl=knl, ksl=ksl, lrad=L / n, l=0.0, tilt=tilt),
                                      el.position + s))
        elif base in ("sbend", "rbend") and L > 0:
            angle = float(getattr(el, "angle", 0.0))
            k1 = float(getattr(el, "k1", 0.0))
            h = angle / L
            e1, e2 = float(getattr(el, "e1", 0.0)), float(getattr(el, "e2", 0.0))
            if base == "rbend":
                e1, e2 = e1 + angle / 2, e2 + angle / 2
            hgap, fint = float(getattr(el, "hgap", 0.0)), float(getattr(el, "fint", 0.0))
            out.append(MadElement(el.name + "_den", "dipedge", dict(h=h, e1=e1, hgap=hgap, fint=fint, l=0.0),
                                  el.position))
            for i, s in enumerate(_teapot(L, n)):
                out.append(MadElement("%s..%d" % (el.name, i + 1) if n > 1 else el.name, "multipole",
                                      dict(knl=[angle / n, k1 * L / n], ksl=[0.0, 0.0], lrad=L / n, l=0.0,
                                           tilt=tilt), el.position + s))
            out.append(MadElement(el.name + "_dex", "dipedge", dict(h=h, e1=e2, hgap=hgap, fint=fint, l=0.0),
                                  el.position + L))
        elif base in ("hkicker", "vkicker", "kicker", "tkicker"):
            attrs = dict(lrad=L, l=0.0, tilt=tilt)
            for k in ("kick", "hkick", "vkick"):
                if hasattr(el, k):
                    attrs[k] = float(getattr(el, k))
            out.append(MadElement(el.name, base, attrs, el.position + 0.5 * L))
        elif base == "rfcavity":
            attrs = dict(volt=float(getattr(el, "volt", 0.0)), freq=float(getattr(el, "freq", 0.0)),
                         lag=float(getattr(el, "lag", 0.0)), l=0.0)
            out.append(MadElement(el.name, base, attrs, el.position + 0.5 * L))
        else:
            attrs = {k: v for k, v in vars(el).items() if k not in ("name", "base_type", "position")}
            out.append(MadElement(el.name, base, attrs, el.position))
    out.sort(key=lambda e: e.position)
    return MadSequence(seq.name, seq.length, out)


_DRIFT_LIKE = ("marker", "monitor", "hmonitor", "vmonitor", "collimator", "rcollimator", "elseparator",
               "instrument", "solenoid", "drift")


def iter_from_madx_sequence(sequence, classes, ignored_madtypes=(), exact_drift=False,
                            drift_threshold=1e-6, install_apertures=False):
    """``(name, element)`` pairs for a thin sequence -- restates the mapping of
    ``xline/loader_mad.py:6-249`` (implicit drifts :29-32, element table :40-183, tilt
    wrappers :186-197, apertures :199-246, closing drift :248-249) for the element types this
    package supports."""
    if not isinstance(classes, dict):
        classes = {k: getattr(classes, k) for k in dir(classes) if not k.startswith("_")}
    Drift = classes["DriftExact"] if exact_drift else classes["Drift"]
    Multipole = classes["Multipole"]
    old_pp, i_drift = 0.0, 0
    pairs = sorted(zip(sequence.element_positions(), sequence.elements), key=lambda t: t[0])
    for pp, ee in pairs:
        if pp > old_pp + drift_threshold:
            yield "drift_%d" % i_drift, Drift(length=(pp - old_pp))
            old_pp = pp
            i_drift += 1
        kind = ee.base_type.name
        new = None
        if kind in _DRIFT_LIKE:
            new = Drift(length=ee.l)
            old_pp += ee.l
        elif kind in ignored_madtypes:
            pass
        elif kind == "multipole":
            knl = list(getattr(ee, "knl", [0]))
            ksl = list(getattr(ee, "ksl", [0]))
            new = Multipole(knl=knl, ksl=ksl, hxl=knl[0], hyl=ksl[0], length=ee.lrad)
        elif kind in ("tkicker", "kicker"):
            new = Multipole(knl=[-ee.hkick] if hasattr(ee, "hkick") else [],
                            ksl=[ee.vkick] if hasattr(ee, "vkick") else [], length=ee.lrad, hxl=0, hyl=0)
        elif kind == "vkicker":
            new = Multipole(knl=[], ksl=[ee.kick], length=ee.lrad, hxl=0, hyl=0)
        elif kind == "hkicker":
            new = Multipole(knl=[-ee.kick], ksl=[], length=ee.lrad, hxl=0, hyl=0)
        elif kind == "dipedge":
            new = classes["DipoleEdge"](h=ee.h, e1=ee.e1, hgap=ee.hgap, fint=ee.fint)
        elif kind == "rfcavity":
            new = classes["Cavity"](voltage=ee.volt * 1e6, frequency=ee.freq * 1e6, lag=ee.lag * 360)
        elif kind == "placeholder":
            slot = int(getattr(ee, "slot_id", 0))
            if slot in (1, 2, 3):
                new = classes[{1: "SCCoasting", 2: "SCQGaussProfile", 3: "SCInterpolatedProfile"}[slot]]()
            else:
                new = Drift(length=ee.l)
                old_pp += ee.l
        else:
            raise ValueError('MAD element "%s" not recognized' % kind)
        tilt = math.degrees(ee.tilt) if abs(getattr(ee, "tilt", 0.0)) > 0 else 0
        if abs(tilt) > 0:
            yield ee.name + "_pretilt", classes["SRotation"](angle=tilt)
        yield ee.name, new
        if abs(tilt) > 0:
            yield ee.name + "_posttilt", classes["SRotation"](angle=-tilt)
        if install_apertures and hasattr(ee, "aperture") and min(ee.aperture) > 0:
            ap = ee.aperture
            if ee.apertype == "rectangle":
                yield ee.name + "_aperture", classes["LimitRect"](min_x=-ap[0], max_x=ap[0], min_y=-ap[1], max_y=ap[1])
            elif ee.apertype == "ellipse":
                yield ee.name + "_aperture", classes["LimitEllipse"](a=ap[0], b=ap[1])
            elif ee.apertype == "circle":
                yield ee.name + "_aperture", classes["LimitEllipse"](a=ap[0], b=ap[0])
            elif ee.apertype == "rectellipse":
                yield ee.name + "_aperture", classes["LimitRectEllipse"](max_x=ap[0], max_y=ap[1], a=ap[2], b=ap[3])
            else:
                raise ValueError("Aperture type not recognized")
    if sequence.length > old_pp:
        yield "drift_%d" % i_drift, Drift(length=(sequence.length - old_pp))
