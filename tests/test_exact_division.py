"""The strict kernels replace the IEEE division subroutine by a reciprocal known in advance plus
FMA remainder corrections (xline_b200/csrc/track_impl.cuh: ``div_small_int`` for the ``/ ii`` of
the Horner step, xline/elements.py:130-134; ``div_known_recip`` for ``/ length`` and ``/ (a*a)``,
:143-144, :436).  Bit-identity with the reference rests on those sequences returning the correctly
rounded quotient: the same sequences restated in C (tests/exact_division.c) are compared here with
the IEEE division on random dividends and on dividends constructed next to rounding midpoints of
the quotient; the GPU self-test (``xlb_selftest_exact_division``) repeats it on the device."""
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    out = str(tmp_path_factory.mktemp("exact_division") / "exact_division")
    # -ffp-contract=off: the C compiler must not fuse anything on its own
    subprocess.run([cc, "-O2", "-ffp-contract=off", "-o", out, os.path.join(HERE, "exact_division.c"), "-lm"],
                   check=True)
    return out


def _run(exe, *args):
    res = subprocess.run([exe] + [str(a) for a in args], check=True, capture_output=True, text=True)
    done, bad = (int(v) for v in res.stdout.split())
    return done, bad, res.stderr


def test_division_by_small_integers_is_correctly_rounded(exe):
    done, bad, err = _run(exe, 0, 200_000, 1)
    assert done >= 255 * 200_000 and bad == 0, err


def test_division_by_recorded_reciprocal_is_correctly_rounded(exe):
    done, bad, err = _run(exe, 1, 200_000, 2)
    assert done >= 256 * 200_000 and bad == 0, err


def test_division_by_the_lengths_and_apertures_of_the_shipped_lattices(exe):
    """The divisors the strict kernel actually meets on C2 / C4 / C5: multipole lengths and the
    squared half-axes of the elliptic apertures."""
    from xline_b200 import configs

    divs = set()
    for fn in (configs.config_lhc, configs.config_petra4, configs.config_psb):
        line = fn(8)[0]
        for el in line.elements:
            nm = type(el).__name__
            if nm == "Multipole" and (el.hxl != 0 or el.hyl != 0) and el.length > 0:
                divs.add(float(el.length))
            elif nm in ("LimitEllipse", "LimitRectEllipse"):
                divs.add(float(el.a * el.a))
                divs.add(float(el.b * el.b))
    divs = sorted(divs)
    assert len(divs) > 3
    done, bad, err = _run(exe, 1, 100_000, 3, *[np.format_float_scientific(d, unique=True) for d in divs[:400]])
    assert bad == 0, err


@pytest.mark.gpu
def test_device_division_sequences_match_device_ieee_division():
    from xline_b200 import _cabi

    bad, n = _cabi.selftest_exact_division(list(range(1, 256)), mode=0, samples_per_thread=64, seed=11)
    assert n > 2e9 and bad == 0
    rng = np.random.default_rng(5)
    divs = list(np.exp(rng.uniform(-40, 40, 60))) + [0.1, 3.0, 14.3, 0.022 ** 2, 0.018 ** 2, 1.0 - 2 ** -53]
    bad, n = _cabi.selftest_exact_division(divs, mode=1, samples_per_thread=128, seed=12)
    assert n > 1e9 and bad == 0
    # wide exponent range (the tracking kernels guard |a| < 2^-959 themselves)
    bad, n = _cabi.selftest_exact_division([3, 7, 11, 13, 14], mode=0, samples_per_thread=256, seed=13,
                                           exponent_span=900)
    assert bad == 0
