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


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "hostemu")], check=True)
    lib = ctypes.CDLL(os.path.join(HERE, "hostemu", "_build", "libhostemu.so"))
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


def test_reference_flow_headers_reproduce_the_oracle(emu, golden):
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


def test_symphony_diagnostics_headers_match_the_oracle(emu, oracle):
    d = oracle.make_dist(oracle.PITCHY_PL, [2.5, 1.0])
    pv = (ctypes.c_double * 2)(2.5, 1.0)
    s, theta = 50.0, 0.9
    for what, a, b in ((0, 90.0, 2.2), (1, 90.0, 0.0), (2, 80.0, 200.0), (3, 2.0, 0.0), (3, 300.0, 0.0)):
        for coeff in (0, 1):
            for stokes in (0, 1, 2):
                want = oracle.symphony_diagnostic(d, coeff, stokes, s, theta, what, a, b)
                got = emu.lib.emu_symphony_diag(oracle.PITCHY_PL, pv, 2, coeff, stokes, s, theta, what, a, b)
                assert abs(got / want - 1) < 1e-9, (what, coeff, stokes, got, want)
