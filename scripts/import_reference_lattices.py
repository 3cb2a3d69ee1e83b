"""Converts the SixTrack lattices shipped with the reference (``/root/reference/examples``)
into this package's own serialised ``Line`` format (gzip JSON of ``Line.to_dict()``), so
that the benchmark configurations of BASELINE.json can be rebuilt on the GPU box, where
the reference tree does not exist.  Runs in the build container only.

    python scripts/import_reference_lattices.py
"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xline_b200.line import Line  # noqa: E402
from xline_b200.sixtrack_input import SixInput  # noqa: E402

REF = os.environ.get("XLINE_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "xline_b200", "lattices")


def convert(example, out_name, synth_fort16=False):
    six = SixInput(os.path.join(REF, "examples", example))
    if synth_fort16:
        six.synthesize_fort16(seed=20261018)
    line = Line.from_sixinput(six)
    d = line.to_dict(keepextra=True)
    e0 = six.initialconditions[-1] * 1e6
    d["meta"] = dict(
        source="examples/%s (SixTrack fort.2/fort.3/fort.8)" % example,
        energy0_eV=e0, mass0_eV=six.pma * 1e6, harm=six.harm, tlen=six.tlen,
        synthetic_fort16=bool(synth_fort16), n_elements=len(line),
    )
    fn = os.path.join(OUT, out_name + ".json.gz")
    with gzip.GzipFile(fn, "wb", mtime=0) as fh:
        fh.write(json.dumps(d, separators=(",", ":")).encode())
    kinds = {}
    for el in line.elements:
        kinds[type(el).__name__] = kinds.get(type(el).__name__, 0) + 1
    print(out_name, len(line), kinds, "%.1f kB" % (os.path.getsize(fn) / 1e3))


def convert_madx(example_file, seq_name, out_name, slices, energy0_eV, mass0_eV):
    from xline_b200.madx_input import MadxFile, makethin

    mf = MadxFile(os.path.join(REF, example_file))
    thin = makethin(mf.sequence(seq_name), slices)
    line = Line.from_madx_sequence(thin, exact_drift=True)
    d = line.to_dict(keepextra=True)
    d["meta"] = dict(source="%s (MAD-X sequence %s, TEAPOT thin slicing %s, exact drifts)"
                     % (example_file, seq_name, slices),
                     energy0_eV=energy0_eV, mass0_eV=mass0_eV, tlen=thin.length, n_elements=len(line))
    fn = os.path.join(OUT, out_name + ".json.gz")
    with gzip.GzipFile(fn, "wb", mtime=0) as fh:
        fh.write(json.dumps(d, separators=(",", ":")).encode())
    print(out_name, len(line), "%.1f kB" % (os.path.getsize(fn) / 1e3))


def convert_psb():
    """tests/psb/psb_fb_lhc.madx executed as the reference's test does
    (tests/test_madx_import.py:40-52): call the sequence / aperture / strength files, flatten,
    ``makethin`` with one TEAPOT slice and dipedges, ring ``psb1``, apertures installed.  The
    files leave ``kBHZ`` (the main-bend angle knob, psb.seq:1777-1808) undefined -- it lives in
    an orbit file that is not shipped -- so MAD-X would track a ring without bends; the value
    written in the magnet class itself (psb.seq:92, ANGLE = -TWOPI/32) is used.  With it the
    shipped kQF / kQD give QH = 4.22, QV = 4.45, the targets of the MATCH block."""
    import math

    from xline_b200 import optics
    from xline_b200.madx_input import MadxFile

    mad = MadxFile(os.path.join(REF, "tests/psb/psb_fb_lhc.madx"), defaults={"kbhz": -2 * math.pi / 32})
    seq = mad.sequence["psb1"]
    line = Line.from_madx_sequence(seq, install_apertures=True)
    tw = optics.twiss(line)
    pc = mad.beam["pc"] * 1e9
    m0 = 938.27208816e6
    d = line.to_dict(keepextra=True)
    d["meta"] = dict(source="tests/psb/psb_fb_lhc.madx (MAD-X sequence psb1, makethin slice=1 teapot, "
                            "makedipedge, apertures installed; kbhz = -twopi/32)",
                     energy0_eV=math.sqrt(pc * pc + m0 * m0), mass0_eV=m0, tlen=seq.length,
                     n_elements=len(line), qx=tw["qx"], qy=tw["qy"], skipped=mad.skipped)
    fn = os.path.join(OUT, "psb.json.gz")
    with gzip.GzipFile(fn, "wb", mtime=0) as fh:
        fh.write(json.dumps(d, separators=(",", ":")).encode())
    print("psb", len(line), "Q = %.6f %.6f" % (tw["qx"], tw["qy"]), "%.1f kB" % (os.path.getsize(fn) / 1e3))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    convert_psb()
    # examples/petra4/track_p1.py:21-30: 6 GeV electrons, 4 slices for sbend and quadrupole
    convert_madx("examples/petra4/h7ba_n8.seq", "ring", "petra4", {"sbend": 4, "quadrupole": 4},
                 6e9, 0.51099895e6)
    convert("fodo", "fodo")
    convert("lhc", "lhc", synth_fort16=True)
    convert("bbsimple", "bbsimple")
    convert("beambeam", "lhc_beambeam")
