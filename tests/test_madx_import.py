"""MAD-X sequence import without cpymad (xline_b200/madx_input.py): parser, thin slicing,
and the element mapping pinned against the reference's own loader."""
import math
import os

import numpy as np
import pytest

import xline_b200 as xl
from xline_b200.madx_input import MadxFile, iter_from_madx_sequence, makethin

SMALL = """
! a toy ring in the MAD-X subset the reader supports
lcell = 10;  kf := 0.3 * scale; scale = 1.0;   // deferred expression, defined before use
kd := -kf;
qf: quadrupole, l := 0.4, k1 := kf;
qd: qf, k1 := kd;                               // class inheritance overrides k1 only
b1: sbend, l = 1.5, angle := twopi / 8, e1 := twopi/16, e2 := twopi/16, hgap = 0.02, fint = 0.5;
sx: sextupole, l = 0.2, k2 = 1.5;
bpm: monitor, l = 0.1;
cav: rfcavity, l = 0, volt := 2 * 1.5, freq = 500, lag = 0.5;
kh: hkicker, l = 0, kick = 1e-4;
ring: sequence, l = 4 * lcell;
qf, at = 0.2;   sx, at = 0.6;  b1, at = 2.5;  bpm, at = 4.0;  qd, at = 5.2;  b1, at = 7.5;  kh, at = 9.0;
qf, at = 10.2;  b1, at = 12.5; qd, at = 15.2; b1, at = 17.5; cav, at = 19;
qf, at = 20.2;  b1, at = 22.5; qd, at = 25.2; b1, at = 27.5;
qf, at = 30.2;  b1, at = 32.5; qd, at = 35.2; b1, at = 37.5;
endsequence;
"""


def test_parser_expressions_inheritance_and_positions():
    mf = MadxFile(text=SMALL)
    assert mf.value("kd") == -0.3 and mf.value("twopi/8") == pytest.approx(math.pi / 4)
    base, attrs = mf.element_attrs("qd")
    assert base == "quadrupole" and attrs["k1"] == -0.3 and attrs["l"] == 0.4
    seq = mf.sequence("ring")
    assert seq.length == 40.0 and len(seq.elements) == 20
    assert seq.elements[0].position == pytest.approx(0.0)      # at = centre, entry = at - l/2
    assert seq.elements[2].position == pytest.approx(2.5 - 0.75)


def test_makethin_teapot_and_mapping():
    seq = MadxFile(text=SMALL).sequence("ring")
    thin = makethin(seq, {"quadrupole": 4, "sbend": 2})
    line = xl.Line.from_madx_sequence(thin)
    assert line.get_length() == pytest.approx(40.0, abs=1e-12)
    kinds = [type(e).__name__ for e in line.elements]
    assert kinds.count("DipoleEdge") == 16 and kinds.count("Cavity") == 1
    quads = [e for e in line.elements if isinstance(e, xl.Multipole) and len(e.knl) == 2 and e.hxl == 0
             and e.knl[1] != 0]
    assert len(quads) == 8 * 4 and quads[0].knl[1] == pytest.approx(0.3 * 0.4 / 4)
    bends = [e for e in line.elements if isinstance(e, xl.Multipole) and e.hxl != 0]
    assert sum(b.hxl for b in bends) == pytest.approx(2 * math.pi)
    # TEAPOT: first kick of a 4-slice quad at L/10 from the entry, inner spacing 4L/15
    s = line.get_s_elements()
    i0 = [i for i, e in enumerate(line.elements) if e is quads[0]][0]  # Element.__eq__ is by value
    i1 = [i for i, e in enumerate(line.elements) if e is quads[1]][0]
    assert s[i0] == pytest.approx(0.04) and s[i1] - s[i0] == pytest.approx(0.4 * 4 / 15)
    cav = [e for e in line.elements if isinstance(e, xl.Cavity)][0]
    assert (cav.voltage, cav.frequency, cav.lag) == (3e6, 5e8, 180.0)
    kick = [e for e in line.elements if isinstance(e, xl.Multipole) and e.knl == [-1e-4]]
    assert len(kick) == 1
    with pytest.raises(ValueError):
        list(iter_from_madx_sequence(seq, xl.elements.element_classes()))  # thick quadrupole: not recognised


def test_shipped_petra4_lattice():
    from xline_b200 import configs

    line, meta = configs.load_lattice("petra4")
    kinds = {}
    for e in line.elements:
        kinds[type(e).__name__] = kinds.get(type(e).__name__, 0) + 1
    assert kinds["Cavity"] == 2 and kinds["DipoleEdge"] == 2 * 1540  # h7ba_n8.seq: 1540 sbend placements
    assert line.get_length() == pytest.approx(meta["tlen"], abs=1e-9)
    assert sum(e.hxl for e in line.elements if isinstance(e, xl.Multipole)) == pytest.approx(2 * math.pi, rel=1e-8)


@pytest.mark.skipif(not os.path.isdir("/root/reference/xline"), reason="reference tree absent")
def test_mapping_against_reference_loader():
    """The reference's own iter_from_madx_sequence, fed with this reader's thin sequence
    objects, must yield the same element list as the restated mapping."""
    from oracle import ref_harness as rh

    els = rh.load_reference()
    import importlib

    ref_loader = importlib.import_module("xline.loader_mad") if False else None
    import sys, types
    pkg = sys.modules.get("xline")
    if pkg is None or not hasattr(pkg, "__path__"):
        stub = types.ModuleType("xline")
        stub.__path__ = [os.path.join(rh.REFERENCE_ROOT, "xline")]
        sys.modules["xline"] = stub
    ref_loader = importlib.import_module("xline.loader_mad")
    for text, slices in ((SMALL, {"quadrupole": 3, "sbend": 2}),):
        thin = makethin(MadxFile(text=text).sequence("ring"), slices)
        for exact in (False, True):
            mine = list(iter_from_madx_sequence(thin, xl.elements.element_classes(), exact_drift=exact))
            theirs = list(ref_loader.iter_from_madx_sequence(thin, classes=els, exact_drift=exact))
            assert [n for n, _ in mine] == [n for n, _ in theirs]
            for (_, a), (_, b) in zip(mine, theirs):
                assert type(a).__name__ == type(b).__name__
                da, db = a.to_dict(), b.to_dict()
                for k in da:
                    if k != "__class__":
                        assert np.array_equal(np.asarray(da[k], dtype=float), np.asarray(db[k], dtype=float)), k
    petra = makethin(MadxFile(os.path.join(rh.REFERENCE_ROOT, "examples/petra4/h7ba_n8.seq")).sequence("ring"),
                     {"sbend": 4, "quadrupole": 4})
    mine = list(iter_from_madx_sequence(petra, xl.elements.element_classes(), exact_drift=True))
    theirs = list(ref_loader.iter_from_madx_sequence(petra, classes=els, exact_drift=True))
    assert len(mine) == len(theirs) == 31025
    assert all(type(a).__name__ == type(b).__name__ and a.to_dict().keys() == b.to_dict().keys()
               for (_, a), (_, b) in zip(mine, theirs))
