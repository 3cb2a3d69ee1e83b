"""CPU-side tests: element API, lattice packing, SixTrack reader, C-ABI surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import xline_b200 as xl
from xline_b200 import _cabi, lattice
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_element_constructors_like_reference():
    """reference tests/test_elements.py:4-32 and tests/test_line.py:11 (positional)."""
    for cls in (xl.XYShift, xl.SRotation, xl.Cavity, xl.Line, xl.DipoleEdge):
        cls()
    assert xl.Drift(length=4).length == 4 and xl.DriftExact(length=4).length == 4
    assert xl.Drift(0).length == 0
    el = xl.Multipole()
    assert el.order == 0 and el.knl[0] == 0 and el.ksl[0] == 0
    el = xl.Multipole(knl=[1])
    assert el.knl == [1] and el.ksl == [0] and el.order == 0
    assert xl.Multipole(knl=[1, 2, 3]).order == 2 and xl.Multipole(ksl=[1, 2, 3]).order == 2
    a, b = xl.Multipole(), xl.Multipole()
    a.knl.append(3)
    assert b.knl == [0]  # list defaults are per-instance
    with pytest.raises(TypeError):
        xl.Drift(foo=1)


def test_element_dict_roundtrip_and_extra_fields():
    bb = xl.BeamBeam4D(charge=1e11, sigma_x=1e-3, sigma_y=2e-3)
    assert bb.min_sigma_diff == 1e-28 and bb.enabled is True
    assert "enabled" not in bb.to_dict() and "enabled" in bb.to_dict(keepextra=True)
    assert xl.BeamBeam4D.from_dict(bb.to_dict(keepextra=True)) == bb
    assert bb.copy() == bb and bb.copy() is not bb
    sc = xl.SCQGaussProfile()
    assert sc.get_fields() == ["number_of_particles", "bunchlength_rms", "sigma_x", "sigma_y",
                               "length", "x_co", "y_co"]
    assert sc.q_parameter == 1.0


def test_line_editing_and_serialisation(tmp_path):
    line = xl.Line([xl.Drift(1.0), xl.Drift(0.0), xl.Multipole(knl=[0, 0]), xl.Drift(2.0),
                    xl.Multipole(knl=[0, 0.1]), xl.Cavity(voltage=1e6, frequency=4e8)])
    assert len(line) == 6 and line.get_length() == 3.0
    assert len(line.remove_zero_length_drifts()) == 5
    assert len(line.remove_inactive_multipoles()) == 5
    merged = line.remove_zero_length_drifts().remove_inactive_multipoles().merge_consecutive_drifts()
    assert len(merged) == 3 and merged.elements[0].length == 3.0
    fn = str(tmp_path / "line.json")
    line.to_json(fn)
    back = xl.Line.from_json(fn)
    assert back.to_dict() == line.to_dict()
    assert line.get_s_elements() == [0.0, 1.0, 1.0, 1.0, 3.0, 3.0]


def test_particles_reference_quantities():
    """reference tests/test_particles.py:9-25."""
    p = xl.Particles(p0c=1e9, device="cpu")
    for setter, val in (("beta0", 0.91), ("beta0", 0.9101), ("gamma0", 1.99), ("p0c", 0.1 * p.mass0)):
        setattr(p, setter, val)
        err = abs(p.p0c ** 2 + p.mass0 ** 2 - p.energy0 ** 2) / p.mass0 ** 2
        assert err < 1.2e-15


def test_particles_loss_compaction_cpu():
    """reference tests/test_losses.py:5-17."""
    import torch

    p = xl.Particles(p0c=1e9, x=np.arange(10, dtype=np.float64), device="cpu")
    p.state = (p.x.to(torch.int64) % 2 == 0).to(torch.int64)
    p.remove_lost_particles()
    p.state = (p.x > 5).to(torch.int64)
    p.remove_lost_particles()
    assert p.x.tolist() == [6.0, 8.0]
    assert [len(lp) for lp in p.lost_particles] == [5, 3]


def test_particles_delta_setter_matches_oracle():
    from oracle import xline_oracle as xo

    d = np.linspace(-1e-3, 1e-3, 7)
    p = xl.Particles(p0c=450e9, delta=d, device="cpu")
    o = xo.OracleParticles(7, p0c=450e9, delta=d)
    assert np.array_equal(p.rpp.numpy(), o.rpp) and np.array_equal(p.rvv.numpy(), o.rvv)


def _walk(packed):
    """Decode the packed words back into (tag, aux, element_index) triples."""
    out = []
    w = packed.words
    for c in range(packed.n_chunks):
        pos = c * packed.chunk_words
        while True:
            hdr = int(w[pos])
            tag, aux, size, idx = hdr & 0xFF, (hdr >> 8) & 0xFF, (hdr >> 16) & 0x3FFF, hdr >> 32
            if tag in (lattice.T_END_CHUNK, lattice.T_END_TURN):
                break
            out.append((tag, aux, idx))
            pos += 2 * size
    return out


def test_pack_structure_and_validation():
    line = xl.Line([xl.Drift(1.0), xl.Drift(0.0), xl.Multipole(knl=[0, 0.1, 3.0], ksl=[0, 0, 0, 6.0]),
                    xl.Multipole(knl=[1e-3], hxl=1e-3, length=2.0), xl.LimitRect(), xl.Cavity(voltage=1.0),
                    xl.RFMultipole(knl=[1, 2]), xl.SRotation(angle=3), xl.XYShift(dx=1),
                    xl.DipoleEdge(h=1), xl.LimitEllipse(), xl.LimitRectEllipse(), xl.DriftExact(1.0),
                    xl.BeamMonitor(num_stores=2, max_particle_id=9)])
    line.fuse_records = False
    pk = line.pack()
    recs = _walk(pk)
    assert [r[2] for r in recs] == [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]  # zero drift dropped
    assert recs[1][:2] == (lattice.T_MULTIPOLE, 3)
    assert recs[2][0] == lattice.T_MULTIPOLE_CURVED
    assert pk.monitor_words == 7 * 2 * 10
    # fast encoding folds 1/i!
    w = pk.words.view(np.float64)
    base = 2  # after the drift record
    assert w[base + 2] == 0.0 and w[base + 3] == 6.0 / 6.0
    assert w[base + 4] == 3.0 / 2.0
    ws = line.pack(strict=True).words.view(np.float64)
    assert ws[base + 3] == 6.0 and ws[base + 4] == 3.0
    assert line.pack(strict=True).flags & lattice.F_STRICT
    # corrupt a header: the C-side validator must reject it
    bad = pk.words.copy()
    bad[0] = np.uint64(77)
    lat = _cabi.Lattice(bad.ctypes.data, bad.size, pk.chunk_words, pk.n_chunks, pk.n_elements, pk.flags)
    assert _cabi.lib().xlb_lattice_validate(C.byref(lat)) != 0
    assert b"unknown tag" in _cabi.lib().xlb_last_error()


def test_pack_fuses_multipole_aperture_drift():
    line = xl.Line([xl.Multipole(knl=[0, 0.1]), xl.LimitEllipse(a=1, b=2), xl.Drift(3.0),
                    xl.Multipole(knl=[1e-3], hxl=1e-3, length=2.0), xl.Drift(0.0), xl.Drift(4.0),
                    xl.LimitRect(), xl.Multipole(knl=[0, 0, 1]), xl.LimitRect(min_x=-2, max_x=1)])
    pk = line.pack()
    recs = _walk(pk)
    L = lattice
    assert [(r[0], r[2]) for r in recs] == [
        (L.T_THIN_BLOCK | L.AP_ELLIPSE | L.TB_DRIFT, 0), (L.T_THIN_BLOCK | L.TB_CURVED | L.TB_DRIFT, 3),
        (L.T_LIMIT_RECT, 6), (L.T_THIN_BLOCK | L.AP_RECT, 7)]
    w = pk.words
    f = pk.words.view(np.float64)
    assert f[1] == 3.0 and int(w[2]) == 1  # drift length, aperture element index
    size0 = (int(w[0]) >> 16) & 0x3FFF
    assert f[2 * size0 + 1] == 4.0
    line.fuse_records = False
    assert len(_walk(line.pack())) == 8
    # a symmetric rectangle right after a multipole uses the |x| <= max form (fast encoding only)
    sym = xl.Line([xl.Multipole(knl=[0, 1]), xl.LimitRect(min_x=-2, max_x=2, min_y=-1, max_y=1)])
    assert _walk(sym.pack())[0][0] == L.T_THIN_BLOCK | L.AP_RECT_SYM
    assert _walk(sym.pack(strict=True))[0][0] == L.T_THIN_BLOCK | L.AP_RECT


def test_pack_chunking_never_splits_a_record():
    els = []
    for i in range(3000):
        els += [xl.Drift(1.0 + i), xl.Multipole(knl=list(range(1, 16)), ksl=[0.0] * 15)]
    pk = xl.Line(els).pack()
    assert pk.n_chunks > 10
    recs = _walk(pk)
    # leading drift, then (multipole -> drift) pairs fused, last multipole alone
    assert len(recs) == 3001 and [r[2] for r in recs] == [0] + list(range(1, 6000, 2))
    big = xl.SCInterpolatedProfile(number_of_particles=1.0, line_density_profile=list(np.ones(3000)),
                                   sigma_x=1.0, sigma_y=2.0, length=1.0)
    pk = xl.Line([xl.Drift(1.0), big]).pack()
    assert pk.chunk_words >= 3000 and pk.flags & lattice.F_BEAMFIELDS


def test_pack_error_behaviour():
    with pytest.raises(ZeroDivisionError):  # gaussian_fields.py:91-92 semantics at pack time
        xl.Line([xl.SCCoasting(number_of_particles=1.0, sigma_x=1.0, sigma_y=1.0, min_sigma_diff=0.0)]).pack()
    xl.Line([xl.BeamBeam4D(charge=1.0, sigma_x=1.0, sigma_y=1.0)]).pack()  # round branch
    class LimitPolygon(xl.Element):
        _base = ()
    with pytest.raises(NotImplementedError):
        xl.Line([LimitPolygon()]).pack()


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "xline_b200.h")).read()
    declared = set(re.findall(r"\b(xlb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_cabi.EXPORTS)
    L = _cabi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.xlb_abi_version() == _cabi.ABI_VERSION == 4
    # argument validation happens before any CUDA call
    assert L.xlb_track_device(None, None, None, None) == -1
    assert b"null" in L.xlb_last_error()


def test_sixtrack_reader_builds_shipped_lattices():
    from xline_b200 import configs

    line, meta = configs.load_lattice("lhc")
    kinds = {}
    for el in line.elements:
        kinds[type(el).__name__] = kinds.get(type(el).__name__, 0) + 1
    # 8 316 Drift + 10 137 Multipole + 12 Cavity (+ fort.8 wrappers)
    assert kinds["Drift"] == 8316 and kinds["Multipole"] == 10137 and kinds["Cavity"] == 12
    assert abs(line.get_length() - meta["tlen"]) < 1e-6
    steps = sum(el.order for el in line.elements if isinstance(el, xl.Multipole))
    assert steps == 73394  # Horner steps per turn
    fodo, _ = configs.load_lattice("fodo")
    assert [type(e).__name__ for e in fodo.elements].count("Cavity") == 1
    bb, _ = configs.load_lattice("lhc_beambeam")
    assert sum(isinstance(e, xl.BeamBeam4D) for e in bb.elements) == 72
    assert sum(isinstance(e, xl.BeamBeam6D) for e in bb.elements) == 2


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="reference tree absent")
def test_sixtrack_reader_against_reference_loader():
    """expand_struct vs the reference's own _expand_struct fed with this SixInput."""
    from oracle import ref_harness as rh
    from xline_b200.sixtrack_input import SixInput, expand_struct

    els = rh.load_reference()
    ref_loader = rh.reference_module("loader_sixtrack")
    for ex in ("fodo", "lhc", "bbsimple"):
        six = SixInput("/root/reference/examples/" + ex)
        mine, _, iconv = expand_struct(six, xl.elements.element_classes())
        six2 = SixInput("/root/reference/examples/" + ex)
        theirs, _, iconv2 = ref_loader._expand_struct(six2, convert=els)
        assert iconv == iconv2
        assert [(n, t) for n, t, _ in mine] == [(n, t) for n, t, _ in theirs]
        for (_, _, a), (_, _, b) in zip(mine, theirs):
            da, db = a.to_dict(), b.to_dict()
            for k in da:
                if k != "__class__":
                    assert np.array_equal(np.asarray(da[k], dtype=float), np.asarray(db[k], dtype=float)), k


def test_algorithmic_ops_match_oracle_count():
    from oracle import xline_oracle as xo
    from xline_b200 import configs

    line, _ = configs.load_lattice("lhc")
    assert line.algorithmic_ops_per_turn() == xo.algorithmic_ops(line.to_specs())
    assert abs(line.algorithmic_ops_per_turn() - 9.48e5) / 9.48e5 < 0.02


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under xline_b200/ or scripts/ may import or
    execute it (only tests/, __graft_entry__.smoke() and bench.py's CPU legs do)."""
    for top in ("xline_b200", "scripts"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                    txt = open(os.path.join(dirpath, fn)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                    assert "xline_oracle" not in txt and "ref_harness" not in txt, fn
                    assert "run_oracle" not in txt, fn


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", os.path.join(ROOT, "xline_b200", "does_not_exist.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _cabi.lib()


def test_packed_lattice_binary_roundtrip(tmp_path):
    line = xl.Line([xl.Drift(1.0), xl.Multipole(knl=[0, 0.1, 3.0]), xl.LimitRect(), xl.Drift(2.0),
                    xl.BeamMonitor(num_stores=2, max_particle_id=9), xl.Cavity(voltage=1e6, frequency=4e8)])
    pk = line.pack()
    fn = str(tmp_path / "lattice.npz")
    pk.save(fn)
    back = lattice.PackedLattice.load(fn)
    assert np.array_equal(back.words, pk.words) and back.words.dtype == np.uint64
    assert (back.chunk_words, back.n_chunks, back.n_elements, back.flags) == (
        pk.chunk_words, pk.n_chunks, pk.n_elements, pk.flags)
    assert back.monitor_layout == pk.monitor_layout and back.record_counts == pk.record_counts
    lat = _cabi.Lattice(back.words.ctypes.data, back.words.size, back.chunk_words, back.n_chunks,
                        back.n_elements, back.flags)
    assert _cabi.lib().xlb_lattice_validate(C.byref(lat)) == 0


def test_particles_derived_longitudinal_variables():
    d = np.array([-1e-3, 0.0, 2e-3])
    p = xl.Particles(p0c=450e9, delta=d, zeta=[0.1, -0.2, 0.3], device="cpu")
    e = p.energy.numpy()
    pc = p.pc.numpy()
    assert np.allclose(e ** 2, pc ** 2 + p.mass0 ** 2, rtol=1e-14)       # E^2 = (pc)^2 + m^2
    assert np.allclose(p.ptau.numpy(), (e - p.energy0) / p.p0c, rtol=0, atol=1e-15)
    beta = pc / e
    assert np.allclose(p.rvv.numpy(), beta / p.beta0, rtol=1e-14)          # rvv = beta / beta0
    assert np.allclose(p.tau.numpy() * p.beta0, p.sigma.numpy(), rtol=1e-15)
    assert np.allclose(p.psigma.numpy() * p.beta0, p.ptau.numpy(), rtol=1e-15)
    assert np.allclose(p.mass_ratio.numpy(), 1.0)


def test_line_like_reference_test_line():
    """reference tests/test_line.py:7-94, line for line."""
    rfmultipole = xl.RFMultipole(frequency=100, knl=[0.1, 0.2], ksl=[0.3, 0.4])
    zero_drift = xl.Drift(0)
    line = xl.Line(elements=[zero_drift, rfmultipole, zero_drift],
                   element_names=["zero_drift", "rfmultipole", "zero_drift"])
    length = 1.4
    n_elements, position = 3, 1
    line.insert_element(position, xl.DriftExact(length), "exact drift")
    n_elements += 1
    assert len(line) == n_elements and line.find_element_ids("exact drift")[0] == position
    assert line.get_length() == length
    line.insert_element(position, xl.Multipole(knl=[0.1]), "multipole")
    line.insert_element(position + 1, xl.LimitEllipse(a=0.08, b=0.02), "multipole_aperture")
    n_elements += 2
    assert len(line) == n_elements
    for kw in (dict(dx=0, dy=0), dict(dx=0.2, dy=-0.003)):
        line._add_offset_error_to("multipole", **kw)
        n_elements += 2
        assert len(line) == n_elements
    for angle in (0, 0.1):
        line._add_tilt_error_to("multipole", angle=angle)
        n_elements += 2
        assert len(line) == n_elements
    line._add_multipole_error_to("multipole", knl=[0, 0.1], ksl=[-0.03, 0.01])
    mp = line.elements[line.element_names.index("multipole")]
    assert mp.knl == [0.1, 0.1] and mp.ksl == [-0.03, 0.01]
    # wrappers sit outside the element AND its aperture
    i0, i1 = line.find_element_ids("multipole")
    assert line.element_names[i0 - 1] == "multipole_tilt_in" and line.element_names[i1] == "multipole_tilt_out"
    line_dict = line.to_dict()
    line = xl.Line.from_dict(line_dict)
    assert len(line) == n_elements
    line.append_line(xl.Line.from_dict(line_dict))
    n_elements *= 2
    assert len(line) == n_elements and line.get_length() == 2 * length
    sd, su = line.get_s_elements("downstream"), line.get_s_elements("upstream")
    assert max(np.array(sd) - np.array(su)) == length
    line.insert_element(1, xl.Multipole(), "inactive_multipole")
    n_elements += 1
    assert len(line.remove_inactive_multipoles()) == n_elements - 1 and len(line) == n_elements
    line.remove_inactive_multipoles(inplace=True)
    n_elements -= 1
    assert len(line) == n_elements
    assert len(line.merge_consecutive_drifts()) == n_elements - 1
    line.merge_consecutive_drifts(inplace=True)
    n_elements -= 1
    assert len(line) == n_elements
    assert len(line.get_elements_of_type(xl.Drift)) == 2
    drifts = line.get_elements_of_type(xl.Drift)[0]
    nz = len([d for d in drifts if d.length == 0])
    assert len(line.remove_zero_length_drifts()) == n_elements - nz
    line.remove_zero_length_drifts(inplace=True)
    assert len(line) == n_elements - nz
    # reference tests/test_line.py:97-105 (isthick attribute)
    thick = xl.Multipole(knl=[0, -1.0], ksl=[0, 0], length=4)
    thick.isthick = True
    line2 = xl.Line(elements=[xl.Drift(length=1.0), xl.Multipole(knl=[0, 1.0], ksl=[0, 0]), xl.Drift(length=3), thick])
    assert np.isclose(line2.get_length(), 8.0, rtol=1e-30, atol=1e-20)
    merged = xl.Line([xl.Multipole(knl=[0, 1.0]), xl.Multipole(knl=[1e-3, 0.5, 2.0], ksl=[0, 0.1])]).merge_consecutive_multipoles()
    assert len(merged) == 1 and merged.elements[0].knl == [1e-3, 1.5, 2.0] and merged.elements[0].ksl == [0, 0.1, 0]
    assert line2.get_element_ids_of_type(xl.Drift, start_idx_offset=2) == [2, 4]


def _bb6d(**kw):
    base = dict(phi=1e-4, alpha=0.3, x_bb_co=1e-5, y_bb_co=-2e-5, charge_slices=[1e10, 2e10, 1e10],
                zeta_slices=[0.05, 0.0, -0.05], sigma_11=1e-9, sigma_12=1e-12, sigma_13=0.0, sigma_14=0.0,
                sigma_22=1e-11, sigma_23=0.0, sigma_24=0.0, sigma_33=2e-9, sigma_34=-1e-12, sigma_44=3e-11)
    base.update(kw)
    return xl.BeamBeam6D(**base)


def test_pack_segments_lattices_with_6d_lenses(tmp_path):
    """Fast encoding: every BeamBeam6D record sits alone in a chunk of its own, the lattice is
    a sequence of segments (xlb_lattice_t::segments) and the closing segment is a tracking
    segment; strict / element-by-element encodings stay one piece."""
    els = [xl.Drift(1.0), xl.Multipole(knl=[0, 0.1]), _bb6d(), xl.LimitRect(min_x=-1, max_x=1, min_y=-1, max_y=1),
           xl.Drift(2.0), _bb6d(phi=2e-4), _bb6d(phi=3e-4), xl.BeamBeam4D(charge=1e10, sigma_x=1e-4, sigma_y=2e-4),
           xl.Drift(1.0), _bb6d(phi=4e-4)]
    line = xl.Line(els)
    pk = line.pack()
    S, M, B = pk.segments.tolist(), lattice.SEG_MAIN, lattice.SEG_BB6D
    assert [s[2] for s in S] == [M, B, M, B, B, M, B, M]
    assert [s[0] for s in S] == list(range(8)) and all(s[1] == 1 for s in S) and pk.n_chunks == 8
    assert pk.flags & lattice.F_BB6D and pk.flags & lattice.F_BEAMFIELDS       # the BeamBeam4D
    assert not (xl.Line(els[:7]).pack().flags & lattice.F_BEAMFIELDS)          # 6D lenses only: lean kernels
    for first, _, kind in S:
        hdr = int(pk.words[first * pk.chunk_words])
        assert ((hdr & 0xff) == lattice.T_BEAMBEAM6D) == (kind == B)
    # element indices stay those of the Line
    lens_idx = [int(pk.words[s[0] * pk.chunk_words]) >> 32 for s in S if s[2] == B]
    assert lens_idx == [2, 5, 6, 9]
    assert line.pack(strict=True).segments is None and (line.pack(strict=True).flags & lattice.F_BEAMFIELDS)
    line.split_lenses = False
    whole = line.pack()
    assert whole.segments is None and whole.n_chunks == 1 and whole.flags & lattice.F_BEAMFIELDS
    line.split_lenses = True
    # on-disk round trip keeps the segment table
    fn = str(tmp_path / "seg.npz")
    pk.save(fn)
    back = lattice.PackedLattice.load(fn)
    assert np.array_equal(back.segments, pk.segments) and back.segments.dtype == np.int32
    L = _cabi.lib()
    lat = back.c_lattice()
    assert L.xlb_lattice_validate(C.byref(lat)) == 0
    # the validator knows the rules
    seg = pk.segments.copy()
    seg[1, 2] = M                                   # a 6D record inside a tracking segment
    lat = pk.c_lattice()
    lat.segments = seg.ctypes.data
    assert L.xlb_lattice_validate(C.byref(lat)) != 0 and b"MAIN segment" in L.xlb_last_error()
    seg = pk.segments.copy()
    seg[2, 0] = 3                                   # hole in the tiling
    lat.segments = seg.ctypes.data
    assert L.xlb_lattice_validate(C.byref(lat)) != 0 and b"tile" in L.xlb_last_error()
    seg = pk.segments[:-1].copy()                   # must end with a tracking segment / cover all chunks
    lat.segments, lat.n_segments = seg.ctypes.data, len(seg)
    assert L.xlb_lattice_validate(C.byref(lat)) != 0
    lat = pk.c_lattice()
    lat.flags = pk.flags & ~lattice.F_BB6D
    assert L.xlb_lattice_validate(C.byref(lat)) != 0 and b"XLB_F_BB6D" in L.xlb_last_error()


def test_edit_tracking_repacks_only_when_something_changed():
    """element fields edited in place must reach the packed lattice; an untouched
    line must not be re-packed (the check is O(1) while the global edit clock stands still)."""
    from xline_b200 import elements as E

    line = xl.Line([xl.Multipole(knl=[0.0, 1e-3]), xl.Drift(length=1.0), xl.Cavity(voltage=1e6, frequency=4e8)])
    a = line.pack()
    assert line.pack() is a and line.pack(strict=True) is line.pack(strict=True)
    clock = E.edit_clock()
    assert line.pack() is a and E.edit_clock() == clock
    xl.Drift(length=3.0)  # an unrelated element is created: clock moves, this line is unchanged
    assert line.pack() is a
    line.elements[2].voltage = 2e6
    b = line.pack()
    assert b is not a and not np.array_equal(a.words, b.words)
    line.elements[0].knl[1] = 2e-3
    c = line.pack()
    assert c is not b and not np.array_equal(c.words, b.words)
    line.elements[0].ksl = np.array([0.0, 1e-4])
    d = line.pack()
    line.elements[0].ksl *= 2.0
    e = line.pack()
    assert d is not c and e is not d and not np.array_equal(d.words, e.words)
    line.elements.append(xl.Drift(length=0.5))
    f = line.pack()
    assert f is not e and f.n_elements == 4
    line.elements += [xl.Drift(length=0.25)]
    assert line.pack().n_elements == 5
    line.fuse_records = False
    assert line.pack() is not f
    # element values own their data: the caller's list is copied on assignment
    mine = [0.0, 5e-3]
    m = xl.Multipole(knl=mine)
    mine[1] = 7.0
    assert m.knl[1] == 5e-3 and m.to_dict()["knl"] == [0.0, 5e-3] and type(m.to_dict()["knl"]) is list
    assert m.copy() == m and xl.Multipole.from_dict(m.to_dict()) == m
