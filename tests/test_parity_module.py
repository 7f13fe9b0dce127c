"""Host logic added in round 2, without a GPU: the parity statistics (rimphony_b200/parity.py), the
prefix-stable sampler, the CPU sample of bench.py, and the reference-divergence rule of the Heyvaerts
product path as compiled for the host (tests/hostemu, a development harness)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import rimphony_b200 as R
from rimphony_b200 import parity as P

HERE = os.path.dirname(os.path.abspath(__file__))


def test_reference_defined_mask_and_stats():
    base = np.ones((8, 6))
    base[6] = [1.0, np.nan, 2.0, np.nan, 3.0, 4.0]
    tight = base.copy()
    tight[6] = [1.0005, np.nan, 2.1, 5.0, np.nan, 4.0]      # ok, NaN on both, moved 5 %, NaN flipped twice, ok
    nudge = base.copy()
    nudge[6, 5] = 4.0 * (1 + 2e-3)                           # moved by 2e-3 under the nudge
    d = P.reference_defined(base, None, (tight, nudge), mask=0xC0)
    assert d[6].tolist() == [True, True, False, False, False, False]
    assert d[:6].all() and d[7].all()
    got = base.copy()
    got[6] = [1.0 + 5e-4, np.nan, 9.0, np.nan, np.nan, 4.0]
    got[7, 0] = np.nan
    got[0, 1] = -1.0
    st = P.parity_stats(got, base, None, d)
    q = st["rho_Q"]
    assert (q["undefined"], q["nan_both"], q["nan_mismatch"], q["finite"], q["within"]) == (4, 1, 0, 1, 1)
    assert st["rho_V"]["nan_mismatch"] == 1 and st["j_I"]["sign_mismatch"] == 1
    assert not P.meets_north_star(st)
    s = P.summarize(st)
    assert s["slots"][6] == "rho_Q" and s["reference_undefined"][6] == 4 and s["n"] == 6


def test_stokes_v_is_measured_on_the_lobe_scale():
    want = np.ones((8, 3))
    want[4] = 1e-3                      # the two lobes (1 and -0.999) nearly cancel
    lobes = np.array([[1.0] * 3, [-0.999] * 3, [1.0] * 3, [1.0] * 3])
    got = want.copy()
    got[4] = 1e-3 + 1e-3                # 100 % of the sum, 5e-4 of the lobe scale
    st = P.parity_stats(got, want, lobes, None)
    assert st["j_V"]["within"] == 3 and st["j_V"]["max_err"] == pytest.approx(1e-3 / 1.999)


def test_fixture_with_stability_companion_loads():
    fx = P.load_fixture("pitchy_pl_4k")
    assert fx["defined"].shape == fx["out"].shape and fx["stability_mask"] == 0xC0
    frac = 1 - fx["defined"][6:].mean()
    assert 0.005 < frac < 0.05          # 1.7-2.4 % of the rho entries (tests/golden/heyvaerts_low_s.md)
    # j / alpha: only the handful of entries where the reference's own error exceeds 1e-3 (make_converged.py)
    assert fx["converged_entries"] > 0 and (~fx["defined"][:6]).sum() <= 0.001 * fx["defined"][:6].size


def test_sampler_prefix_is_independent_of_the_batch_size():
    k1, s1, t1, p1 = R.synthetic_batch("pitchy_kappa", 100, seed=3, shard=2)
    k2, s2, t2, p2 = R.synthetic_batch("pitchy_kappa", 5000, seed=3, shard=2)
    assert np.array_equal(s1, s2[:100]) and np.array_equal(t1, t2[:100])
    assert all(np.array_equal(a, b[:100]) for a, b in zip(p1[:3], p2[:3]))
    # the 1e4-point parity fixture is the prefix of rank 0's benchmark batch
    fx = np.load(os.path.join(HERE, "golden", "pitchy_pl_10k.npz"))
    import bench
    _, s, th, par = bench.draw("pitchy_pl", 20000)
    assert np.array_equal(fx["s"], s[:10000]) and np.array_equal(fx["theta"], th[:10000])
    assert np.array_equal(fx["params"][1], par[1][:10000])


def test_cpu_sample_slices_are_distinct_and_seeded():
    import bench
    a = bench.cpu_sample("pitchy_pl", 0, 8)
    b = bench.cpu_sample("pitchy_pl", 8, 8)
    c = bench.cpu_sample("pitchy_pl", 0, 16)
    assert np.array_equal(np.concatenate([a[1], b[1]]), c[1]) and not np.array_equal(a[1], b[1])
    kj, sj, tj, pj = bench.cpu_sample("juettner_sweep", 0, 64)
    assert kj == R.THERMAL_JUETTNER and len(np.unique(np.round(np.log10(sj), 6))) > 30   # strides through the grid
    assert bench.coeff_mask_of("juettner_sweep") == 0xC0 and bench.coeff_mask_of("pitchy_pl") == 0xFF


def test_kernel_constants_file_is_complete():
    import bench
    kc = bench.kernel_constants()
    for key in ("flop_per_application", "dram_bytes_per_point"):
        assert set(kc[key]) == {"symphony", "heyvaerts"} and all(v > 0 for v in kc[key].values())
    assert "source" in kc


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "hostemu")], check=True)
    lib = ctypes.CDLL(os.path.join(HERE, "hostemu", "_build", "libhostemu.so"))
    dp, up = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint)
    lib.emu_point.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                              dp, dp, dp, up]
    return lib


def _emu_rho(lib, kind, params, s, theta):
    pv = (ctypes.c_double * len(params))(*params)
    eps = (ctypes.c_double * 4)(1e-3, 1e-3, 1e-3, 1e-3)
    out, lob, info = (ctypes.c_double * 8)(), (ctypes.c_double * 4)(), (ctypes.c_uint * 4)()
    assert lib.emu_point(kind, pv, len(params), 2, 2, s, theta, eps, out, lob, info) == 0
    return out[6], out[7], info[2], info[3]


def test_reference_divergence_rule_on_the_host_build(emu_lib):
    """rb_heyfast.cuh kHeyRefDiverges*: the power laws with gamma_min = 1 return NaN for rho_V below
    s = 0.38 (0.33 isotropic) and for rho_Q below a threshold that rises with theta (0.18 ... 0.37; 0 ... 0.30
    isotropic), with STATUS_REFERENCE_DIVERGES and no rule application where both are NaN; other kinds, and a
    power law with gamma_min > 1, integrate."""
    pl = [2.5, 1.0, 1.0, 1e12, 1e10]
    q, v, apps, status = _emu_rho(emu_lib, R.PITCHY_PL, pl, 0.2, 0.9)
    assert np.isnan(q) and np.isnan(v) and apps == 0 and status & R.STATUS_REFERENCE_DIVERGES
    q, v, apps, status = _emu_rho(emu_lib, R.PITCHY_PL, pl, 0.375, 0.9)
    assert np.isfinite(q) and np.isnan(v) and apps > 0 and status & R.STATUS_REFERENCE_DIVERGES
    q, v, apps, status = _emu_rho(emu_lib, R.PITCHY_PL, pl, 0.25, 0.1)     # small angle: rho_Q is computed
    assert np.isfinite(q) and np.isnan(v) and apps > 0
    q, v, apps, status = _emu_rho(emu_lib, R.POWER_LAW, [2.5, 1.0, 1e12, 1e10], 0.25, 0.3)
    assert np.isfinite(q) and np.isnan(v) and status & R.STATUS_REFERENCE_DIVERGES
    q, v, apps, status = _emu_rho(emu_lib, R.POWER_LAW, [2.5, 1.0, 1e12, 1e10], 0.25, 1.4)
    assert np.isnan(q) and np.isnan(v) and apps == 0
    q, v, apps, status = _emu_rho(emu_lib, R.PITCHY_PL, pl, 0.7, 0.9)
    assert np.isfinite(q) and np.isfinite(v) and not status & R.STATUS_REFERENCE_DIVERGES
    q, v, apps, status = _emu_rho(emu_lib, R.POWER_LAW, [2.5, 1.5, 1e12, 1e10], 0.2, 0.9)
    assert apps > 0 and not status & R.STATUS_REFERENCE_DIVERGES
    q, v, apps, status = _emu_rho(emu_lib, R.THERMAL_JUETTNER, [10.0], 0.2, 0.9)
    assert apps > 0 and not status & R.STATUS_REFERENCE_DIVERGES


def test_low_s_rule_agrees_with_the_oracle_fixture(emu_lib, golden):
    """The documented agreement of the rule (tests/golden/heyvaerts_low_s.md): on the s < 1 points of the
    4096-point fixture the NaN verdicts agree on >= 85 %, and where both sides are finite the values agree."""
    fx = P.load_fixture("pitchy_pl_4k")
    idx = np.where(fx["s"] < 1.0)[0][:400]
    got = np.full((8, len(idx)), np.nan)
    for j, i in enumerate(idx):
        q, v, _, _ = _emu_rho(emu_lib, R.PITCHY_PL, [float(p[i]) for p in fx["params"]], float(fx["s"][i]), float(fx["theta"][i]))
        got[6, j], got[7, j] = q, v
    st = P.parity_stats(got, fx["out"][:, idx], None, fx["defined"][:, idx], mask=0xC0)
    for nm in ("rho_Q", "rho_V"):
        assert st[nm]["nan_mismatch"] <= 0.15 * len(idx), st[nm]
        assert st[nm]["frac_within"] >= 0.97 and st[nm]["sign_mismatch"] == 0, st[nm]


def test_chunk_growth_vote_is_unanimous_on_the_benchmark_configurations(emu_lib):
    """the product path grows delta_n when all active accumulators vote for it, the
    reference per coefficient (symphony.rs:243-258).  Measured on seeded points of the two crank-out
    configurations: the accumulators never disagree (the vote is |dG/dn / chunk| < 1e-5, decided by the
    common J_n^2 envelope), so the two rules produce the same chunk sequence."""
    emu_lib.emu_vote_stats.argtypes = [ctypes.POINTER(ctypes.c_long)]
    before = (ctypes.c_long * 2)()
    emu_lib.emu_vote_stats(before)
    dp = ctypes.POINTER(ctypes.c_double)
    for config, n in (("pitchy_pl", 160), ("pitchy_kappa", 40)):
        kind, s, th, params = R.synthetic_batch(config, n, seed=11)
        for i in range(n):
            if config == "pitchy_kappa" and (s[i] > 1e4 or params[0][i] < 2.6):
                continue   # hard spectra at high s take the faithful continuation: minutes on the host
            pv = [float(np.asarray(p)[i]) if np.ndim(p) else float(p) for p in params]
            arr = (ctypes.c_double * len(pv))(*pv)
            eps = (ctypes.c_double * 4)(1e-3, 1e-3, 1e-3, 1e-3)
            out, lob, info = (ctypes.c_double * 8)(), (ctypes.c_double * 4)(), (ctypes.c_uint * 4)()
            assert emu_lib.emu_point(kind, arr, len(pv), 2, 1, float(s[i]), float(th[i]), eps, out, lob, info) == 0
    after = (ctypes.c_long * 2)()
    emu_lib.emu_vote_stats(after)
    assert after[0] - before[0] > 1000            # votes were taken ...
    assert after[1] - before[1] == 0              # ... and none of them was split
