"""The algorithmic logic of the CUDA kernels, exercised without a GPU.

tests/hostemu compiles the device headers of rimphony_b200/csrc with g++ and -DRB_HOST_EMU (the lane dimension
of the warp-cooperative quadrature becomes an explicit loop).  It is a DEVELOPMENT HARNESS: the package never
loads it, it is not a fallback, and the parity tests proper (-m gpu) go through the real C ABI on a B200.  What
it gives the CPU suite is an early warning: the same source the kernels are built from, run on a handful of
fixture points, must land where the oracle does.  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
MODE_FAITHFUL, MODE_FAST = 0, 2   # the harness's own numbering (hostemu.cpp: `fused`)


def _load(name):
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "hostemu")], check=True)
    lib = ctypes.CDLL(os.path.join(HERE, "hostemu", "_build", name))
    dp, up = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint)
    lib.emu_point.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                              dp, dp, dp, up]
    lib.emu_symphony_diag.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double]
    lib.emu_symphony_diag.restype = ctypes.c_double

    def point(kind, params, mode, which, s, theta):
        pv = (ctypes.c_double * len(params))(*params)
        eps = (ctypes.c_double * 4)(1e-3, 1e-3, 1e-3, 1e-3)
        out, lob, info = (ctypes.c_double * 8)(), (ctypes.c_double * 4)(), (ctypes.c_uint * 4)()
        assert lib.emu_point(kind, pv, len(params), mode, which, s, theta, eps, out, lob, info) == 0
        return np.array(out[:]), np.array(lob[:]), list(info)
    point.lib = lib
    return point


@pytest.fixture(scope="module")
def emu():
    """the build with the product path's lean math (rb_core.cuh RB_LEAN_MATH)"""
    return _load("libhostemu.so")


@pytest.fixture(scope="module")
def emu_ieee():
    """the build with IEEE division / libm, like the QUADPACK-faithful Heyvaerts kernels"""
    return _load("libhostemu_ieee.so")


def test_product_path_headers_land_on_the_oracle(emu, golden):
    fx = golden("pitchy_pl")
    sigma0 = fx["s"] * np.sin(fx["theta"])
    picks = [i for i in range(len(sigma0)) if sigma0[i] >= 3.0][:10]
    worst = np.zeros(8)
    for i in picks:
        got, lobes, info = emu(fx["kind"], [float(p[i]) for p in fx["params"]], MODE_FAST, 3, float(fx["s"][i]),
                               float(fx["theta"][i]))
        want = fx["out"][:, i]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        for c in range(8):
            if np.isnan(want[c]):
                continue
            scale = abs(want[c])
            if c in (4, 5):   # Stokes V: the two lobes nearly cancel (symphony.rs:97-107)
                scale = abs(fx["lobes"][2 * (c - 4), i]) + abs(fx["lobes"][2 * (c - 4) + 1, i])
            worst[c] = max(worst[c], abs(got[c] - want[c]) / scale)
        assert 0 < info[0] < 5000 and 0 < info[2] < 20000   # rule applications: Symphony, Heyvaerts
    assert (worst < 3e-3).all(), worst


def test_reference_flow_headers_reproduce_the_oracle(emu_ieee, golden):
    emu = emu_ieee
    fx = golden("pitchy_pl")
    sigma0 = fx["s"] * np.sin(fx["theta"])
    picks = [i for i in range(len(sigma0)) if sigma0[i] >= 3.0 and fx["s"][i] < 300][:3]
    for i in picks:
        got, _, _ = emu(fx["kind"], [float(p[i]) for p in fx["params"]], MODE_FAITHFUL, 3, float(fx["s"][i]),
                        float(fx["theta"][i]))
        want = fx["out"][:, i]
        ok = ~np.isnan(want)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.abs(got[ok] / want[ok] - 1).max() < 1e-9   # same libm on both sides here


def test_symphony_diagnostics_headers_match_the_oracle(emu_ieee, oracle):
    emu = emu_ieee
    d = oracle.make_dist(oracle.PITCHY_PL, [2.5, 1.0])
    pv = (ctypes.c_double * 2)(2.5, 1.0)
    s, theta = 50.0, 0.9
    for what, a, b in ((0, 90.0, 2.2), (1, 90.0, 0.0), (2, 80.0, 200.0), (3, 2.0, 0.0), (3, 300.0, 0.0)):
        for coeff in (0, 1):
            for stokes in (0, 1, 2):
                want = oracle.symphony_diagnostic(d, coeff, stokes, s, theta, what, a, b)
                got = emu.lib.emu_symphony_diag(oracle.PITCHY_PL, pv, 2, coeff, stokes, s, theta, what, a, b)
                assert abs(got / want - 1) < 1e-9, (what, coeff, stokes, got, want)


# --- the building blocks of the instruction-count work (DESIGN.md sections 4, 5.8, 6), one at a time ----------
def _lib(emu):
    lib = emu.lib
    dp = ctypes.POINTER(ctypes.c_double)
    lib.emu_lean_math.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double]
    lib.emu_lean_math.restype = ctypes.c_double
    lib.emu_bessel_pair.argtypes = [ctypes.c_double, ctypes.c_double, dp]
    lib.emu_jy_pair.argtypes = [ctypes.c_double, ctypes.c_double, dp]
    lib.emu_i_thirds.argtypes = [ctypes.c_double, dp]
    return lib


def test_lean_math_is_within_two_ulp_and_keeps_the_special_values(emu):
    """rb_exp / rb_log / rb_rcp / rb_sqrt / rb_div of rb_core.cuh (host twin of the device code: the MUFU seeds are
    emulated to 2^-23): accuracy against numpy's longdouble, and the documented behaviour at 0, inf, NaN."""
    lib = _lib(emu)
    rng = np.random.default_rng(11)
    eps = np.finfo(float).eps

    def ulps(got, want):
        return abs(float((np.longdouble(got) - want) / want)) / eps

    worst = [0.0] * 5
    for _ in range(20000):
        x = float(rng.uniform(-700, 700))
        worst[0] = max(worst[0], ulps(lib.emu_lean_math(0, x, 0.0), np.exp(np.longdouble(x))))
        y = float(np.exp(rng.uniform(-60, 60))) if rng.random() < 0.5 else 1.0 + float(rng.uniform(-0.5, 0.5))
        if y != 1.0:
            worst[1] = max(worst[1], ulps(lib.emu_lean_math(1, y, 0.0), np.log(np.longdouble(y))))
        p, q = float(np.exp(rng.uniform(-300, 300))), float(np.exp(rng.uniform(-300, 300)))
        worst[2] = max(worst[2], ulps(lib.emu_lean_math(2, p, 0.0), 1 / np.longdouble(p)))
        worst[3] = max(worst[3], ulps(lib.emu_lean_math(3, p, 0.0), np.sqrt(np.longdouble(p))))
        worst[4] = max(worst[4], ulps(lib.emu_lean_math(4, p, q), np.longdouble(p) / np.longdouble(q)))
    assert worst[0] <= 1.0 and worst[1] <= 2.0 and worst[2] <= 1.0 and worst[3] <= 1.0 and worst[4] <= 1.5, worst

    inf, nan = float("inf"), float("nan")
    f = lib.emu_lean_math
    assert f(0, inf, 0) == inf and f(0, -inf, 0) == 0.0 and np.isnan(f(0, nan, 0)) and f(0, 710.0, 0) == inf
    assert f(0, -745.0, 0) == 5e-324 and f(0, -746.0, 0) == 0.0           # down to the subnormals, like exp()
    assert f(1, 0.0, 0) == -inf and np.isnan(f(1, -1.0, 0)) and f(1, inf, 0) == inf and f(1, 1.0, 0) == 0.0
    assert f(1, 5e-324, 0) == pytest.approx(np.log(5e-324), rel=1e-15)    # subnormal arguments: the library function
    assert np.isnan(f(2, 0.0, 0)) and np.isnan(f(2, inf, 0)) and np.isnan(f(2, 1e-310, 0))  # documented: NaN, not inf / 0
    assert f(3, 0.0, 0) == 0.0 and f(3, 1e-310, 0) == 0.0 and np.isnan(f(3, -1.0, 0)) and f(3, 4.0, 0) == 2.0


def test_paired_bessel_evaluations_equal_the_single_ones(emu):
    """J_n and J_{n+1} from one Miller recurrence (n + 1 < 30) and from the shared-coefficient Debye expansion
    (n >= 30) against pkgw_bessel_j evaluated once per order, over the argument range of the Symphony integrand."""
    lib = _lib(emu)
    rng = np.random.default_rng(12)
    out = (ctypes.c_double * 4)()
    worst = 0.0
    for _ in range(4000):
        small = rng.random() < 0.4
        n = float(rng.integers(0, 29)) if small else float(np.exp(rng.uniform(np.log(30.0), np.log(1e9))))
        # from deep in Meissel's zone to the turning point
        x = n * (1.0 - float(np.exp(rng.uniform(np.log(1e-9), np.log(0.9))))) if n > 0 else 0.0
        if not x > 0.0:
            continue
        lib.emu_bessel_pair(n, x, out)
        for single, pair in ((out[0], out[2]), (out[1], out[3])):
            if single == 0.0 or not np.isfinite(single):
                assert pair == single or (np.isnan(pair) and np.isnan(single))
                continue
            worst = max(worst, abs(pair / single - 1.0))
    assert worst < 1e-11, worst


def test_prepared_jy_orders_equal_the_plain_pair(emu):
    lib = _lib(emu)
    rng = np.random.default_rng(13)
    out = (ctypes.c_double * 8)()
    scipy_special = pytest.importorskip("scipy.special")
    worst, worst_ref = 0.0, 0.0
    for _ in range(4000):
        sigma = float(rng.uniform(1e-3, 3.0))
        x = sigma * float(rng.uniform(1e-4, 1.0))
        lib.emu_jy_pair(sigma, x, out)
        ref = [scipy_special.jv(sigma, x), scipy_special.jv(sigma - 1, x), scipy_special.yv(sigma, x),
               scipy_special.yv(sigma - 1, x)]
        scale_j = max(abs(ref[0]), abs(ref[1]))
        scale_y = max(abs(ref[2]), abs(ref[3]))
        for k in range(4):
            scale = scale_j if k < 2 else scale_y
            worst = max(worst, abs(out[4 + k] - out[k]) / scale)
            worst_ref = max(worst_ref, abs(out[4 + k] - ref[k]) / scale)
    assert worst < 1e-12 and worst_ref < 1e-11, (worst, worst_ref)


def test_i_thirds_give_the_k_like_differences_to_1e6(emu):
    """bessel_i_thirds truncates its series where the differences I_-nu - I_nu = (2/pi) sin(nu pi) K_nu that the
    quasi-resonant elements are built from are good to 3e-7."""
    lib = _lib(emu)
    scipy_special = pytest.importorskip("scipy.special")
    out = (ctypes.c_double * 4)()
    for g in np.concatenate([np.geomspace(1e-6, 1.0, 30), np.linspace(1.0, 9.99, 60)]):
        lib.emu_i_thirds(float(g), out)
        for nu, ip, im in ((1 / 3, out[0], out[1]), (2 / 3, out[2], out[3])):
            assert ip == pytest.approx(scipy_special.iv(nu, g), rel=1e-7)
            k_like = (2 / np.pi) * np.sin(nu * np.pi) * scipy_special.kv(nu, g)
            assert im - ip == pytest.approx(k_like, rel=2e-6), (g, nu)
