"""TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.npz`` by EXECUTING THE REFERENCE
(``/root/reference/xline``) through ``oracle/ref_harness.py``.  Runs only in the build
container; the fixtures it writes are committed so that the GPU box (which has no
``/root/reference``) can check both the oracle and the CUDA path against reference
outputs.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 20261018


def element_cases():
    """(case name, type, fields) -- one or more parameter sets per hot-path element."""
    z15 = np.linspace(-0.15, 0.15, 7)
    cases = [
        ("drift", "Drift", dict(length=2.339)),
        ("drift_exact", "DriftExact", dict(length=5.3)),
        ("multipole_quad", "Multipole", dict(knl=[0.0, 9.6e-4], ksl=[0.0, 0.0])),
        (
            "multipole_ord5_skew",
            "Multipole",
            dict(knl=[1e-5, -2e-3, 0.3, 11.0, -450.0, 9e4], ksl=[0.0, 1e-3, -0.2, 7.0]),
        ),
        (
            "multipole_bend",
            "Multipole",
            dict(knl=[1.887e-4, 0.0, 0.02], ksl=[0.0], hxl=1.887e-4, hyl=0, length=3.4),
        ),
        (
            "multipole_vbend_len0",
            "Multipole",
            dict(knl=[0.0], ksl=[-2.1e-4, 1e-4], hxl=0, hyl=-2.1e-4, length=0),
        ),
        ("cavity", "Cavity", dict(voltage=16e6, frequency=400.79e6, lag=180.0)),
        ("cavity_lag30", "Cavity", dict(voltage=3e6, frequency=200.0e6, lag=30.0)),
        (
            "rfmultipole",
            "RFMultipole",
            dict(
                voltage=1e5,
                frequency=400e6,
                lag=12.0,
                knl=[1e-4, 2e-3, 0.1],
                ksl=[-2e-4, 1e-3],
                pn=[90.0, 10.0, 0.0],
                ps=[0.0, 45.0],
            ),
        ),
        ("rfmultipole_crab", "RFMultipole", dict(frequency=400e6, knl=[3.4e6 / 6.5e12], pn=[90.0])),
        ("xyshift", "XYShift", dict(dx=1.3e-4, dy=-2.4e-4)),
        ("srotation", "SRotation", dict(angle=17.3)),
        ("dipole_edge", "DipoleEdge", dict(h=0.0123, e1=0.05, hgap=0.02, fint=0.5)),
        ("limit_rect", "LimitRect", dict(min_x=-1e-3, max_x=8e-4, min_y=-5e-4, max_y=1.1e-3)),
        ("limit_ellipse", "LimitEllipse", dict(a=1.2e-3, b=7e-4)),
        ("limit_rect_ellipse", "LimitRectEllipse", dict(max_x=1e-3, max_y=5e-4, a=1.2e-3, b=7e-4)),
        (
            "bb4d_ellip",
            "BeamBeam4D",
            dict(charge=1.15e11, sigma_x=1.2e-3, sigma_y=4.0e-4, beta_r=0.98, x_bb=1e-4,
                 y_bb=-2e-4, d_px=1e-9, d_py=-2e-9),
        ),
        (
            "bb4d_ellip_tall",
            "BeamBeam4D",
            dict(charge=1.15e11, sigma_x=3.0e-4, sigma_y=9.0e-4, beta_r=1.0, x_bb=0.0, y_bb=0.0),
        ),
        ("bb4d_round", "BeamBeam4D", dict(charge=2e11, sigma_x=5e-4, sigma_y=5e-4, beta_r=1.0)),
        (
            "sc_coasting",
            "SCCoasting",
            dict(number_of_particles=1e11, circumference=157.0, sigma_x=1.5e-3, sigma_y=0.9e-3,
                 length=1.3, x_co=1e-4, y_co=-1e-4),
        ),
        (
            "sc_qgauss",
            "SCQGaussProfile",
            dict(number_of_particles=1e11, bunchlength_rms=0.22, sigma_x=0.9e-3, sigma_y=1.4e-3,
                 length=2.0, x_co=0.0, y_co=0.0),
        ),
        (
            "sc_qgauss_q12",
            "SCQGaussProfile",
            dict(number_of_particles=1e11, bunchlength_rms=0.3, sigma_x=1e-3, sigma_y=1e-3,
                 length=2.0, q_parameter=1.2),
        ),
        (
            "sc_qgauss_q08",
            "SCQGaussProfile",
            dict(number_of_particles=1e11, bunchlength_rms=0.3, sigma_x=1.1e-3, sigma_y=1e-3,
                 length=2.0, q_parameter=0.8),
        ),
        (
            "sc_interp_lin",
            "SCInterpolatedProfile",
            dict(number_of_particles=1e11, line_density_profile=list(np.exp(-z15 ** 2 / 0.01)),
                 dz=0.05, z0=-0.15, sigma_x=1.5e-3, sigma_y=0.9e-3, length=1.0, method=0),
        ),
        (
            "sc_interp_cubic",
            "SCInterpolatedProfile",
            dict(number_of_particles=1e11, line_density_profile=list(np.exp(-z15 ** 2 / 0.01)),
                 dz=0.05, z0=-0.15, sigma_x=1.5e-3, sigma_y=0.9e-3, length=1.0, method=1),
        ),
        (
            "bb6d",
            "BeamBeam6D",
            dict(phi=1.4e-4, alpha=0.3, x_bb_co=2e-5, y_bb_co=-1e-5,
                 charge_slices=[3e10, 4e10, 3.5e10, 1e10], zeta_slices=[-0.05, 0.07, 0.0, 0.12],
                 sigma_11=2.1e-8, sigma_12=-1e-10, sigma_13=3e-9, sigma_14=1e-11,
                 sigma_22=4e-10, sigma_23=-2e-11, sigma_24=1e-12, sigma_33=1.1e-8,
                 sigma_34=2e-10, sigma_44=5e-10,
                 x_co=1e-6, px_co=2e-7, y_co=-1e-6, py_co=1e-7, zeta_co=1e-4, delta_co=1e-6,
                 d_x=1e-9, d_px=2e-10, d_y=-1e-9, d_py=1e-10, d_zeta=1e-8, d_delta=1e-11),
        ),
        (
            "bb6d_headon_uncoupled",
            "BeamBeam6D",
            dict(phi=0.0, alpha=0.0, charge_slices=[5e10, 6e10], zeta_slices=[0.03, -0.03],
                 sigma_11=1.6e-8, sigma_22=3e-10, sigma_33=0.9e-8, sigma_44=4e-10),
        ),
        (
            "bb6d_round_crossing",
            "BeamBeam6D",
            dict(phi=1.5e-4, alpha=np.pi / 2, charge_slices=[5e10, 6e10, 4e10],
                 zeta_slices=[0.06, 0.0, -0.06],
                 sigma_11=1.6e-8, sigma_22=3e-10, sigma_33=1.6e-8, sigma_44=3e-10),
        ),
    ]
    return cases


def beam(rng, n, scale=1.0):
    return dict(
        x=rng.normal(0, 6e-4 * scale, n),
        px=rng.normal(0, 3e-5 * scale, n),
        y=rng.normal(0, 5e-4 * scale, n),
        py=rng.normal(0, 2e-5 * scale, n),
        zeta=rng.normal(0, 0.08, n),
        delta=rng.normal(0, 3e-4, n),
    )


def to_jsonable(f):
    out = {}
    for k, v in f.items():
        if isinstance(v, (list, tuple, np.ndarray)):
            out[k] = [float(t) for t in v]
        elif isinstance(v, (bool, np.bool_)):
            out[k] = bool(v)
        elif isinstance(v, (int, np.integer)):
            out[k] = int(v)
        else:
            out[k] = float(v)
    return out


def run_reference(specs, cols, p0c, mass0, num_turns=1, chi=None, charge_ratio=None,
                  scalar_loop=False):
    els = rh.load_reference()
    objs = []
    for name, f in specs:
        cls = getattr(els, name)
        obj = cls(**{k: v for k, v in f.items() if k in cls._base_fields})
        for k, v in f.items():
            if k in cls._extra_fields:
                setattr(obj, k, v)
        objs.append(obj)
    kw = dict(cols)
    if chi is not None:
        kw["chi"] = chi
    if charge_ratio is not None:
        kw["charge_ratio"] = charge_ratio
    if scalar_loop:
        # the reference's q != 1 q-Gaussian only works on scalar particles
        # (qgauss.py:31 ``if u_plus < 0``): run one scalar particle at a time
        n = len(kw["x"])
        out = {}
        for i in range(n):
            p = rh.RefParticles(p0c=p0c, mass0=mass0, **{k: float(v[i]) for k, v in kw.items()})
            for el in objs:
                el.track(p)
            for k in "x px y py zeta s".split():
                out.setdefault(k, []).append(float(getattr(p, k)))
            for k in "delta rpp rvv".split():
                out.setdefault(k, []).append(float(getattr(p, "_" + k)))
            out.setdefault("state", []).append(int(p.state))
            out.setdefault("at_element", []).append(0)
            out.setdefault("at_turn", []).append(1)
        out = {k: np.array(v) for k, v in out.items()}
        out["particle_id"] = np.arange(n)
        return out
    p = rh.RefParticles(p0c=p0c, mass0=mass0, **kw)
    return rh.ref_line_track(objs, p, num_turns=num_turns)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    rng = np.random.default_rng(SEED)
    n = 64
    manifest = {}
    for case, name, f in element_cases():
        cols = beam(rng, n)
        if name.startswith("Limit"):
            cols["x"][:4] = [f.get("max_x", f.get("a", 1)), -f.get("max_x", f.get("a", 1)), 0.0, np.nan]
            cols["y"][:4] = [0.0, 0.0, f.get("max_y", f.get("b", 1)), 0.0]
        if case == "bb4d_round":
            cols["x"][0] = cols["y"][0] = 0.0  # linearised branch r2 < 1e-20
        if case == "sc_interp_lin":
            cols["zeta"][:2] = [-0.5, 0.5]  # outside the profile: np.interp clamps
        p0c = 6.5e12 if not name.startswith("SC") else 0.571e9
        mass0 = 938.27208816e6
        chi = rng.uniform(0.9, 1.1, n) if case in ("multipole_ord5_skew", "rfmultipole") else None
        out = run_reference([(name, f)], cols, p0c, mass0, chi=chi, charge_ratio=chi,
                            scalar_loop=f.get("q_parameter", 1.0) != 1.0)
        arrays = {"in_" + k: v for k, v in cols.items()}
        if chi is not None:
            arrays["in_chi"] = chi
            arrays["in_charge_ratio"] = chi
        arrays.update({"out_" + k: v for k, v in out.items()})
        np.savez(os.path.join(GOLDEN, f"element_{case}.npz"), **arrays)
        manifest[case] = dict(type=name, fields=to_jsonable(f), p0c=p0c, mass0=mass0)
        print(f"{case:28s} ok  lost={int((out['state'] == 0).sum())}")

    # a short mixed line, 3 turns, with apertures (at_element / at_turn bookkeeping)
    line = [
        ("Drift", dict(length=1.5)),
        ("Multipole", dict(knl=[0.0, 0.08], ksl=[0.0, 0.0])),
        ("LimitEllipse", dict(a=2.2e-3, b=1.8e-3)),
        ("DriftExact", dict(length=2.5)),
        ("XYShift", dict(dx=1e-4, dy=-5e-5)),
        ("SRotation", dict(angle=3.0)),
        ("Multipole", dict(knl=[0.0, -0.08, 3.0], ksl=[0.0, 0.0, 1.0])),
        ("SRotation", dict(angle=-3.0)),
        ("XYShift", dict(dx=-1e-4, dy=5e-5)),
        ("LimitRect", dict(min_x=-2.0e-3, max_x=2.1e-3, min_y=-1.9e-3, max_y=1.8e-3)),
        ("Drift", dict(length=1.0)),
        ("DipoleEdge", dict(h=0.01, e1=0.02, hgap=0.01, fint=0.4)),
        ("Multipole", dict(knl=[0.01], ksl=[0.0], hxl=0.01, hyl=0.0, length=1.2)),
        ("Cavity", dict(voltage=2e6, frequency=400e6, lag=180.0)),
        ("RFMultipole", dict(voltage=0.0, frequency=400e6, lag=0.0, knl=[1e-6, 1e-3], ksl=[0, 0],
                             pn=[90.0, 0.0], ps=[0.0, 0.0])),
        ("LimitRectEllipse", dict(max_x=2e-3, max_y=2e-3, a=2.4e-3, b=2.4e-3)),
    ]
    cols = beam(rng, 256, scale=1.6)
    out = run_reference(line, cols, 450e9, 938.27208816e6, num_turns=3)
    arrays = {"in_" + k: v for k, v in cols.items()}
    arrays.update({"out_" + k: v for k, v in out.items()})
    np.savez(os.path.join(GOLDEN, "line_mixed_3turns.npz"), **arrays)
    manifest["line_mixed_3turns"] = dict(
        type="Line", elements=[[n_, to_jsonable(f_)] for n_, f_ in line], p0c=450e9,
        mass0=938.27208816e6, num_turns=3,
    )
    print("line_mixed_3turns ok lost=%d" % int((out["state"] == 0).sum()))

    with open(os.path.join(GOLDEN, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
