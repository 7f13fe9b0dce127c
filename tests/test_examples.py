"""The reference's one-shot example programs (examples/all-pitchykappa-cgs.rs,
one-powerlaw-direct.rs, one-powerlaw-normalized.rs, one-pitchypl-normalized.rs)."""
import io
import math
import re

import pytest

import rimphony_b200 as R
from rimphony_b200 import examples


def test_rust_exponent_formats():
    assert examples.rust_e(2.64399749412774e-21) == "2.64399749412774e-21"  # {:e}: shortest round trip
    assert examples.rust_e(0.0) == "0e0" and examples.rust_e(1.0) == "1e0" and examples.rust_e(-3.25e10) == "-3.25e10"
    assert examples.rust_e(0.001) == "1e-3" and examples.rust_e(123.456) == "1.23456e2"
    assert examples.rust_e(1.0, 18) == "1.000000000000000000e0"
    assert examples.rust_e(float("nan"), 18) == "NaN"
    for x in (math.pi * 1e-30, -math.e * 1e17, 5e-324, 1.7976931348623157e308):
        assert float(examples.rust_e(x)) == x and float(examples.rust_e(x, 18)) == x


class _FakeCalc(R.SynchrotronCalculator):
    def compute_dimensionless(self, coeff, stokes, s, theta):
        return (1 + int(coeff)) * 10.0 + int(stokes) + s * 1e-6


class _FakeModule:
    """The package surface the example programs use, with the device call replaced."""
    def __getattr__(self, name):
        return getattr(R, name)

    class _Dist:
        def __init__(self, *a):
            self.args = a

        def gamma_cutoff(self, c):
            return self

        def gamma_limits(self, *a):
            return self

        def full_calculation(self, mode=0):
            return _FakeCalc()

    PitchyKappaDistribution = PowerLawDistribution = PitchyPowerLawDistribution = _Dist


def test_all_pitchykappa_cgs_layout():
    out = io.StringIO()
    assert examples.main(["all-pitchykappa-cgs", "1e9", "100", "1e3", "0.7", "3.5", "5", "1"], module=_FakeModule(), out=out) == 0
    lines = out.getvalue().splitlines()
    labels = ["    j_I", "alpha_I", "    j_Q", "alpha_Q", "    j_V", "alpha_V", "  rho_Q", "  rho_V"]
    assert [ln.split(": ")[0] for ln in lines] == labels  # all-pitchykappa-cgs.rs:100-131
    assert all(re.fullmatch(r"-?\d\.\d{18}e-?\d+", ln.split(": ")[1]) for ln in lines)
    # compute_cgs scaling (lib.rs:163-173): emission x n_e nu, the others x n_e / nu
    nu, b, n_e = 1e9, 100.0, 1e3
    s = nu / (R.ELECTRON_CHARGE * b / (R.TWO_PI * R.MASS_ELECTRON * R.SPEED_LIGHT))
    assert float(lines[0].split(": ")[1]) == pytest.approx((10.0 + s * 1e-6) * n_e * nu, rel=1e-15)
    assert float(lines[7].split(": ")[1]) == pytest.approx((32.0 + s * 1e-6) * n_e / nu, rel=1e-15)


def test_one_point_examples_layout():
    for tool, pattern in (("one-powerlaw-direct", r"Symphony j_I: 2.64399749412774e-21   Ours: \S+\n"),
                          ("one-powerlaw-normalized", r"Inner Symphony: -?0e0   Us: \S+\nOuter Symphony: 0e0   Us: \S+\n"),
                          ("one-pitchypl-normalized", r"-?\d\.\d{18}e-?\d+\n")):
        out = io.StringIO()
        assert examples.main([tool], module=_FakeModule(), out=out) == 0
        assert re.fullmatch(pattern, out.getvalue()), (tool, out.getvalue())


@pytest.mark.gpu
def test_examples_on_the_device(oracle):
    out = io.StringIO()
    examples.main(["one-powerlaw-direct"], out=out)
    ours = float(out.getvalue().split("Ours: ")[1])
    assert abs(ours / 2.64399749412774e-21 - 1) < 0.01  # one-powerlaw-direct.rs:15

    argv = [1e9, 100.0, 1e3, 0.7, 3.5, 5.0, 1.0]
    out = io.StringIO()
    examples.main(["--mode", "faithful", "all-pitchykappa-cgs"] + [repr(a) for a in argv], out=out)
    got = [float(ln.split(": ")[1]) for ln in out.getvalue().splitlines()]
    d = oracle.make_dist(oracle.PITCHY_KAPPA, [3.5, 5.0, 1.0, 100.0])
    slots = [(oracle.EMISSION, oracle.STOKES_I), (oracle.ABSORPTION, oracle.STOKES_I), (oracle.EMISSION, oracle.STOKES_Q),
             (oracle.ABSORPTION, oracle.STOKES_Q), (oracle.EMISSION, oracle.STOKES_V), (oracle.ABSORPTION, oracle.STOKES_V),
             (oracle.FARADAY, oracle.STOKES_Q), (oracle.FARADAY, oracle.STOKES_V)]
    want = [oracle.compute_cgs(d, c, st, *argv[:4]) for c, st in slots]
    for g, w in zip(got, want):
        assert (math.isnan(g) and math.isnan(w)) or abs(g / w - 1) < 1e-5, (got, want)

    out = io.StringIO()
    examples.main(["one-pitchypl-normalized"], out=out)
    d = oracle.make_dist(oracle.PITCHY_PL, [2.7273434060193211, 2.7016346500930695, 1.0, 1e12, 1e10])
    want = oracle.compute_dimensionless(d, oracle.FARADAY, oracle.STOKES_Q, 8.0973407678629616, 7.2687065355210786e-2)
    got = float(out.getvalue())
    assert (math.isnan(got) and math.isnan(want)) or abs(got / want - 1) < 2e-2
