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
    # the PSB job of the reference's own import test (tests/test_madx_import.py:40-52), apertures installed,
    # and the shipped lattice file is what that job gives
    mad = MadxFile(os.path.join(rh.REFERENCE_ROOT, "tests/psb/psb_fb_lhc.madx"), defaults={"kbhz": -2 * math.pi / 32})
    psb = mad.sequence["psb1"]
    mine = list(iter_from_madx_sequence(psb, xl.elements.element_classes(), install_apertures=True))
    theirs = list(ref_loader.iter_from_madx_sequence(psb, classes=els, install_apertures=True))
    assert [n for n, _ in mine] == [n for n, _ in theirs] and len(mine) == 926
    for (_, a), (_, b) in zip(mine, theirs):
        da, db = a.to_dict(), b.to_dict()
        assert type(a).__name__ == type(b).__name__ and da.keys() == db.keys()
        for k in da:
            if k != "__class__":
                assert np.array_equal(np.asarray(da[k], dtype=float), np.asarray(db[k], dtype=float)), k
    from xline_b200 import configs

    shipped, _ = configs.load_lattice("psb")
    assert shipped.element_names == [n for n, _ in mine]
    assert all(a.to_dict() == b.to_dict() for a, (_, b) in zip(shipped.elements, mine))


# ---------------------------------------------------------------------------------------
# The MAD-X language subset of tests/psb/* and of the reference's own tests/test_madx_import.py
NESTED = """
beam, particle=proton, pc = 0.571;
brho := beam->pc * 3.3356;
mq  : quadrupole, l := 0.5;
mb  : sbend, l := 1.6, angle = -twopi/32, e1 = -twopi/64, e2 = -twopi/64;
oct : multipole, l := 0;
bpm : monitor, l := 0;
tdc : monitor, l := 0.3;
corr: hkicker, l := 0;
r.qf1 : mq;  r.qd1 : mq;  r.b1 : mb;  r.b2 : mb;
cell1: sequence, refer = centre, l = 10;
 r.b1                 , at = 2.0, slot_id = 12345;
 r.qf1                , at = 4.0, slot_id = 1;
 r1.bpm1      : bpm   , at = 4.5, slot_id = 2, assembly_id = 7;
 r1.oct1      : oct   , at = 4.5;
 r.qd1                , at = 6.0;
 r1.tdc       : tdc   , at = 1.5, from = r.qd1;
 r.b2                 , at = 8.5;
endsequence;
cell2: sequence, refer = centre, l = 10;
 r1.corr      : corr  , at = 1.0;
endsequence;
ring1: sequence, refer = entry, l = 20;
 cell1, at = 0;
 cell2, at = 10;
endsequence;
r.b1, angle := kbhz;   r.b2, angle := kbhz;
r.qf1, k1 := kqf;      r.qd1, k1 := kqd;
r1.oct1, knl := {0, 0, 0, koct};
r1.corr, kick := kcorr;
r.b1,  apertype=rectellipse, aperture={0.065, 0.031, 0.065, 0.048};
r.qf1, apertype=ellipse, aperture={0.06, 0.03};
r1.bpm1, apertype=circle, aperture={0.05};
r1.tdc, apertype=rectangle, aperture={0.04, 0.02};
kqf = 0.7; kqd = -0.67; koct = 3.0;
r.qd1->k1 = -0.5;
use, sequence = ring1;
seqedit, sequence = ring1; flatten; endedit;
select, flag=makethin, slice=1;
makethin, sequence=ring1, style=teapot, makedipedge=true;
match, sequence=ring1;
 vary, name=kqf, step=1e-4;
 lmdif, calls=10, tolerance=1e-21;
endmatch;
"""


def test_psb_style_language_subset(tmp_path):
    (tmp_path / "sub").mkdir()
    (tmp_path / "sub" / "Ring.seq").write_text(NESTED)
    (tmp_path / "sub" / "extra.str").write_text("kcorr = 2e-4;\nreturn;\nkcorr = 9;\n")
    (tmp_path / "main.madx").write_text("/* header; with a semicolon */\ncall, file = 'sub/Ring.seq';\n"
                                        "call, file=\"sub/extra.str\"; ! trailing comment\n")
    mad = MadxFile(str(tmp_path / "main.madx"))
    assert mad.skipped == ["match"]
    assert mad.beam == {"particle": "proton", "pc": 0.571} and mad.value("brho") == pytest.approx(0.571 * 3.3356)
    assert mad.value("kcorr") == 2e-4                      # RETURN ends the called file
    thick = mad.sequence("ring1")
    assert [e.name for e in thick.elements] == ["r.b1", "r.qf1", "r1.bpm1", "r1.oct1", "r.qd1", "r1.tdc", "r.b2",
                                               "r1.corr"]
    pos = {e.name: e.position for e in thick.elements}
    assert pos["r.b1"] == pytest.approx(2.0 - 0.8) and pos["r1.tdc"] == pytest.approx(7.5 - 0.15)  # from = r.qd1
    assert pos["r1.corr"] == pytest.approx(11.0)           # refer = entry parent: sub-sequence starts at `at`
    seq = mad.sequence["ring1"]                            # cpymad view: thin, $start / $end markers
    assert seq.elements[0].name == "ring1$start" and seq.elements[-1].name == "ring1$end"
    byname = {e.name: e for e in seq.elements}
    assert byname["r.b1"].knl == [0.0, 0.0]                # undefined kbhz -> 0 overrides the class angle
    assert byname["r.qf1"].knl == [0.0, pytest.approx(0.35)] and byname["r.qd1"].knl[1] == pytest.approx(-0.25)
    assert byname["r1.oct1"].knl == [0, 0, 0, 3.0] and byname["r1.corr"].kick == 2e-4
    assert byname["r.b1_den"].e1 == pytest.approx(-math.pi / 32) and byname["r.b1_den"].h == 0.0
    assert (byname["r.b1"].apertype, byname["r1.bpm1"].aperture) == ("rectellipse", [0.05])
    line = xl.Line.from_madx_sequence(seq, install_apertures=True)
    assert line.get_length() == pytest.approx(20.0)
    kinds = [type(e).__name__ for e in line.elements]
    assert (kinds.count("LimitRectEllipse"), kinds.count("LimitEllipse"), kinds.count("LimitRect")) == (1, 2, 1)
    assert kinds.count("DipoleEdge") == 4
    # defaults= pre-sets variables the files leave undefined
    mad2 = MadxFile(str(tmp_path / "main.madx"), defaults={"kbhz": -2 * math.pi / 32})
    assert {e.name: e for e in mad2.sequence["ring1"].elements}["r.b2"].knl[0] == pytest.approx(-2 * math.pi / 32)


ERRORS = """
    MQ1: Quadrupole, K1:=KQ1, L=1.0, apertype=CIRCLE, aperture={0.04};
    MQ2: Quadrupole, K1:=KQ2, L=1.0, apertype=CIRCLE, aperture={0.04};
    MQ3: Quadrupole, K1:=0.0, L=1.0, apertype=CIRCLE, aperture={0.04};
    KQ1 = 0.02;
    KQ2 = -0.02;
    testseq: SEQUENCE, l = 20.0;
        MQ1, at =  5;
        MQ2, at = 12;
        MQ3, at = 18;
    ENDSEQUENCE;
    BEAM, PARTICLE=PROTON, ENERGY=7000.0, EXN=2.2e-6, EYN=2.2e-6;
    USE, SEQUENCE=testseq;
    Select, flag=makethin, pattern="MQ1", slice=2;
    makethin, sequence=testseq;
    use, sequence=testseq;
    select, flag = error, clear;
    select, flag = error, pattern = "MQ1";
    ealign, dx = 0.01, dy = 0.01, arex = 0.02, arey = 0.02;
    select, flag = error, clear;
    select, flag = error, pattern = "MQ2";
    ealign, dx = 0.04, dy = 0.04, dpsi = 0.1;
    select, flag = error, clear;
    select, flag = error, pattern = "MQ3";
    ealign, dx = 0.00, dy = 0.00, arex = 0.00, arey = 0.00, dpsi = 0.00;
    efcomp, DKN = {0.0, 0.0, 0.001, 0.002}, DKS = {0.0, 0.0, 0.003, 0.004, 0.005};
    select, flag = error, full;
"""


def test_error_import():
    """The reference's tests/test_madx_import.py:55-170 with this package's interpreter in
    place of cpymad: same MAD-X input, same expected element count and order."""
    seq = MadxFile(text=ERRORS).sequence.testseq
    line = xl.Line.from_madx_sequence(seq, install_apertures=True, apply_madx_errors=True)
    expected_element_num = (2 + 6 + 3 + 2 + 3 + 2 + 2 * (3 + 1) + 2 + 2 * 3)
    assert len(line) == expected_element_num
    D, X, M, A, S = xl.Drift, xl.XYShift, xl.Multipole, xl.LimitEllipse, xl.SRotation
    expected_order = [D, D, X, M, X, A, X, X, D, X, D, X, A, X, X, D, X, M, X, A, X, X, D, X, S, M, A, S, X, D,
                      M, A, D, D]
    assert [type(e) for e in line.elements] == expected_order
    mq3 = line.elements[line.element_names.index("mq3")]
    assert abs(mq3.knl[2] - 0.001) < 1e-14 and abs(mq3.knl[3] - 0.002) < 1e-14
    assert abs(mq3.ksl[2] - 0.003) < 1e-14 and abs(mq3.ksl[3] - 0.004) < 1e-14 and abs(mq3.ksl[4] - 0.005) < 1e-14
    # USE drops the error tables again (MAD-X re-expands the sequence)
    mad = MadxFile(text=ERRORS + "use, sequence=testseq;")
    assert len(xl.Line.from_madx_sequence(mad.sequence.testseq, install_apertures=True, apply_madx_errors=True)) == 18


def test_zero_errors():
    """tests/test_madx_import.py:409-449: all-zero error tables load without touching the line."""
    mad = MadxFile(text="""
        qd: multipole, knl={0,-0.3};
        qf: multipole, knl={0, 0.3};
        testseq: sequence, l = 1;
            qd, at = 0.3;
            qf, at = 0.6;
        endsequence;
        beam; use, sequence=testseq;
        select, flag=error, pattern=qf;
        efcomp, dkn={0, 0, 0, 0, 0.0, 0.0, 0.0}, dks={0.0, 0.0, 0, 0};
        ealign, dx=0.0, dy=0.0, ds=0.0, dphi=0.0, dtheta=0.0, dpsi=0.0, mrex=0.0, mrey=0.0, mscalx=0.0,
                mscaly=0.0, arex=0.0, arey=0.0;
    """)
    seq = mad.sequence.testseq
    assert seq.elements[2].align_errors is not None and seq.elements[1].align_errors is None
    line = xl.Line.from_madx_sequence(seq, apply_madx_errors=True)
    assert len(line) == 7 and line.elements[line.element_names.index("qf")].knl == [0, 0.3]
    with pytest.raises(NotImplementedError):
        MadxFile(text="q: multipole; s: sequence, l=1; q, at=0.5; endsequence; use, sequence=s;"
                      "select, flag=error, full; efcomp, radius=0.01, order=1, dknr={0, 1e-4};")


def test_shipped_psb_lattice_and_optics():
    """C5 lattice (tests/psb/psb_fb_lhc.madx -> lattices/psb.json.gz).  Known answer: the
    MATCH block of the MAD-X job (psb_fb_lhc.madx:31-52) tuned the shipped kQF / kQD to
    QH = 4.22, QV = 4.45 on the thin lattice; parser + slicing + optics must reproduce them."""
    from xline_b200 import configs, optics

    line, meta = configs.load_lattice("psb")
    kinds = {}
    for e in line.elements:
        kinds[type(e).__name__] = kinds.get(type(e).__name__, 0) + 1
    assert kinds["DipoleEdge"] == 86 and kinds["Cavity"] == 6 and kinds["LimitRectEllipse"] == 41
    assert kinds["LimitEllipse"] == 221 and kinds["LimitRect"] == 2
    assert line.get_length() == pytest.approx(157.08, abs=1e-9)
    assert sum(e.hxl for e in line.elements if isinstance(e, xl.Multipole)) == pytest.approx(-2 * math.pi)
    tw = optics.twiss(line)
    assert tw["qx"] == pytest.approx(4.22, abs=2e-6) and tw["qy"] == pytest.approx(4.45, abs=2e-6)
    assert 4.0 < tw["alfa_c"] ** -0.5 < 4.4                        # PSB transition gamma ~ 4.1
    # the 16-fold periodicity of the ring shows in the lattice functions
    assert tw["betx"].max() < 8 and tw["bety"].max() < 18 and np.abs(tw["dx"]).max() < 1.8
    line5, cols, p0c, m0 = configs.config_psb(n=2000)
    sc = [e for e in line5.elements if isinstance(e, xl.SCQGaussProfile)]
    assert len(sc) == 120 and line5.get_length() == pytest.approx(157.08, abs=1e-9)
    assert all(e.length == pytest.approx(157.08 / 120) and 2e-3 < e.sigma_x < 6e-3 and 2e-3 < e.sigma_y < 7e-3
               for e in sc)
    tw5 = optics.twiss(line5)
    assert tw5["qx"] == pytest.approx(tw["qx"], abs=1e-9)           # cutting drifts does not change the optics
    assert np.std(cols["x"]) == pytest.approx(np.sqrt(tw["betx"][0] * 1.5e-6 / (p0c / m0)
                                                      + (tw["dx"][0] * 1e-3) ** 2), rel=0.08)


def test_optics_thin_fodo_known_answer():
    """Thin-lens FODO cell: sin(mu/2) = L_cell / (4 f); beta_max/min = L (1 +- sin(mu/2)) / sin(mu)."""
    from xline_b200 import optics

    f, half = 2.5, 2.0
    line = xl.Line([xl.Multipole(knl=[0, 0.5 / f]), xl.Drift(length=half), xl.Multipole(knl=[0, -1 / f]),
                    xl.Drift(length=half), xl.Multipole(knl=[0, 0.5 / f])])
    tw = optics.twiss(line)
    smu2 = half / (2 * f)
    mu = 2 * math.asin(smu2)
    assert tw["qx"] == pytest.approx(mu / (2 * math.pi), abs=1e-12) and tw["qy"] == pytest.approx(tw["qx"], abs=1e-12)
    assert tw["betx"][0] == pytest.approx(2 * half * (1 + smu2) / math.sin(mu), rel=1e-12)
    assert tw["bety"][0] == pytest.approx(2 * half * (1 - smu2) / math.sin(mu), rel=1e-12)
    with pytest.raises(ValueError):
        optics.twiss(xl.Line([xl.Multipole(knl=[0, 2.0]), xl.Drift(length=4.0), xl.Multipole(knl=[0, -2.0]),
                              xl.Drift(length=4.0)]))
    # match_tunes: recover the strengths from the tunes
    def build(kf, kd):
        return xl.Line([xl.Multipole(knl=[0, kf / 2]), xl.Drift(length=half), xl.Multipole(knl=[0, kd]),
                        xl.Drift(length=half), xl.Multipole(knl=[0, kf / 2])])
    kf, kd = optics.match_tunes(build, (0.35, -0.35), tw["qx"], tw["qy"])
    assert kf == pytest.approx(1 / f, rel=1e-8) and kd == pytest.approx(-1 / f, rel=1e-8)


def _hl_like_sequence():
    """Hand-built thin sequence with the element kinds of an HL-LHC job: RF multipole, crab
    cavities (horizontal and vertical), beam-beam markers (4D, 6D) and an octagon aperture."""
    from types import SimpleNamespace

    from xline_b200.madx_input import MadElement, MadSequence

    els = [
        MadElement("rfm", "rfmultipole", dict(volt=1.5, freq=400.0, lag=0.25, knl=[0.0, 1e-3], ksl=[0.0, 2e-4],
                                              pnl=[0.0, 0.1], psl=[0.0, 0.2]), 1.0),
        MadElement("crab_h", "crabcavity", dict(volt=3.4, freq=400.79, lag=0.0, tilt=0.0), 2.0),
        MadElement("crab_v", "crabcavity", dict(volt=3.4, freq=400.79, lag=0.5, tilt=math.pi / 2), 3.0),
        MadElement("bb_ho", "beambeam", dict(slot_id=6), 4.0),
        MadElement("bb_lr", "beambeam", dict(slot_id=4), 5.0),
        MadElement("bb_ho60", "beambeam", dict(slot_id=60), 5.5),
        MadElement("mq", "multipole", dict(knl=[0.0, 1e-3], ksl=[0.0, 0.0], lrad=0.5, apertype="octagon",
                                           aperture=[0.02, 0.018, 0.5, 1.0]), 6.0),
        MadElement("skipme", "instrument", dict(l=0.0), 7.0),
    ]
    return MadSequence("hl", 8.0, els, SimpleNamespace(pc=7000.0))


def test_rf_crab_beambeam_octagon_mappings():
    """xline/loader_mad.py:97-170, 229-242 (RFMultipole, crab cavity with skiptilt, beam-beam
    placeholders by slot id, octagon -> LimitPolygon)."""
    seq = _hl_like_sequence()
    out = dict(iter_from_madx_sequence(seq, xl.elements.element_classes(), install_apertures=True))
    rfm = out["rfm"]
    assert isinstance(rfm, xl.RFMultipole) and rfm.voltage == 1.5e6 and rfm.frequency == 400e6 and rfm.lag == 90.0
    assert list(rfm.pn) == [0.0, 36.0] and list(rfm.ps) == [0.0, 72.0]
    h, v = out["crab_h"], out["crab_v"]
    assert list(h.knl) == [3.4 / 7000.0 * 1e-3] and list(h.pn) == [90.0]
    assert list(v.ksl) == [-3.4 / 7000.0 * 1e-3] and list(v.ps) == [0.5 * 360 + 90]
    assert "crab_v_pretilt" not in out  # skiptilt
    assert isinstance(out["bb_ho"], xl.BeamBeam6D) and isinstance(out["bb_ho60"], xl.BeamBeam6D)
    assert isinstance(out["bb_lr"], xl.BeamBeam4D)
    poly = out["mq_aperture"]
    assert type(poly).__name__ == "LimitPolygon" and len(poly.x_vertices) == 8
    assert poly.x_vertices[0] == 0.02 and poly.y_vertices[1] == 0.018
    # an ignored type is skipped, not yielded as None
    names = [n for n, _ in iter_from_madx_sequence(seq, xl.elements.element_classes(),
                                                   ignored_madtypes=["beambeam"])]
    assert "bb_ho" not in names and "rfm" in names
    # a line holding a LimitPolygon refuses to track like the reference (elements.py:483)
    with pytest.raises(NotImplementedError):
        xl.Line([poly]).pack()
    # Elens: parameters only, readable from the reference's dictionaries
    d = {"elements": [{"__class__": "Elens", "voltage": 1e4, "current": 5.0, "inner_radius": 1e-3,
                       "outer_radius": 2e-3, "ebeam_center_x": 0.0, "ebeam_center_y": 0.0, "elens_length": 3.0}],
         "element_names": ["el"]}
    ln = xl.Line.from_dict(d)
    assert ln.elements[0].current == 5.0
    with pytest.raises(ValueError):
        ln.pack()


@pytest.mark.skipif(not os.path.isdir("/root/reference/xline"), reason="reference tree absent")
def test_rf_crab_beambeam_octagon_mappings_against_reference_loader():
    import importlib
    import sys
    import types

    from oracle import ref_harness as rh

    els = rh.load_reference()
    pkg = sys.modules.get("xline")
    if pkg is None or not hasattr(pkg, "__path__"):
        stub = types.ModuleType("xline")
        stub.__path__ = [os.path.join(rh.REFERENCE_ROOT, "xline")]
        sys.modules["xline"] = stub
    ref_loader = importlib.import_module("xline.loader_mad")
    seq = _hl_like_sequence()
    mine = list(iter_from_madx_sequence(seq, xl.elements.element_classes(), install_apertures=True))
    theirs = list(ref_loader.iter_from_madx_sequence(seq, classes=els, install_apertures=True))
    assert [n for n, _ in mine] == [n for n, _ in theirs]
    for (_, a), (_, b) in zip(mine, theirs):
        assert type(a).__name__ == type(b).__name__
        da, db = a.to_dict(keepextra=False), b.to_dict(keepextra=False)
        assert set(da) == set(db), (type(a).__name__, set(da) ^ set(db))
        for k in da:
            if k != "__class__":
                assert np.array_equal(np.asarray(da[k], dtype=float), np.asarray(db[k], dtype=float)), k
