"""Why the reference returns NaN for rho_Q / rho_V at small s, and how stable that verdict is.
(TEST INFRASTRUCTURE: runs the CPU oracle.  Results: tests/golden/heyvaerts_low_s.md.)

    python tests/golden/study_heyvaerts_low_s.py map [FIXTURE ...]     NaN fraction of the oracle in the (s, theta) plane
    python tests/golden/study_heyvaerts_low_s.py tolerance [N]         the s < 3 points again at epsrel 1e-4 and 1e-5
    python tests/golden/study_heyvaerts_low_s.py nudge [N]             ... and at s (1 +- 1e-4), s (1 + 1e-3)
    python tests/golden/study_heyvaerts_low_s.py where [N]             stage / QAG status / first non-finite element of the failures
    python tests/golden/study_heyvaerts_low_s.py zone I [I ...]        the QR outer integrand F(sigma) at sigma = s (1 +- 10^-j)
    python tests/golden/study_heyvaerts_low_s.py trace I [STOKES]      every bisection of the outer QAG calls of point I

Points are taken from tests/golden/pitchy_pl_4k.npz (I = index into it).
"""
import ctypes
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as O  # noqa: E402


def fixture(name="pitchy_pl_4k"):
    return np.load(os.path.join(HERE, name + ".npz"))


def cmd_map(names):
    names = names or ["pitchy_pl_4k", "pitchy_pl_10k"]
    fxs = [fixture(n) for n in names]
    s = np.concatenate([f["s"] for f in fxs])
    th = np.concatenate([f["theta"] for f in fxs])
    out = np.concatenate([f["out"] for f in fxs], axis=1)
    sb = [0.07, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4, 0.45, 0.5, 0.6, 0.8, 1.0, 1.5, 3.0, 1e4]
    tb = [0, 0.1, 0.2, 0.4, 0.8, 1.2, 1.5708]
    for c, nm in ((6, "rho_Q"), (7, "rho_V")):
        print(f"{nm}: oracle NaN / points, rows s, columns theta  ({', '.join(names)})")
        print("              " + " ".join(f"{tb[j]:.1f}-{tb[j + 1]:.1f} " for j in range(len(tb) - 1)) + "  all")
        for i in range(len(sb) - 1):
            ms = (s >= sb[i]) & (s < sb[i + 1])
            row = []
            for j in range(len(tb) - 1):
                m = ms & (th >= tb[j]) & (th < tb[j + 1])
                row.append(f"{np.isnan(out[c][m]).sum():3d}/{m.sum():<3d}")
            print(f"s {sb[i]:5.2f}-{sb[i + 1]:<7.5g}" + " ".join(row) + f"  {np.isnan(out[c][ms]).mean():.3f}")


def rerun(n, transform=None, epsrel=0.0):
    fx = fixture()
    s, th = fx["s"], fx["theta"]
    sel = np.where(s < 3)[0][:n]
    O.set_epsrel(0.0, epsrel)
    ss = s[sel] if transform is None else transform(s[sel])
    out, _ = O.batch(int(fx["kind"]), ss, th[sel], [p[sel] for p in fx["params"]], coeff_mask=0xC0)
    O.set_epsrel(0.0, 0.0)
    return fx["out"][:, sel], out


def report(label, base, out):
    for c, nm in ((6, "rho_Q"), (7, "rho_V")):
        b, a = base[c], out[c]
        fin = np.isfinite(a) & np.isfinite(b)
        rel = np.abs(a[fin] / b[fin] - 1)
        print(f"{label} {nm}: base NaN {np.isnan(b).sum()} of {len(b)}; finite -> NaN {(np.isfinite(b) & np.isnan(a)).sum()}, "
              f"NaN -> finite {(np.isnan(b) & np.isfinite(a)).sum()}; finite pairs moved > 1e-3: {(rel > 1e-3).sum()} "
              f"(median move {np.median(rel):.1e})", flush=True)


def cmd_tolerance(n):
    for eps in (1e-4, 1e-5):
        base, out = rerun(n, epsrel=eps)
        report(f"epsrel {eps:g}", base, out)


def cmd_nudge(n):
    for nud in (1e-4, -1e-4, 1e-3):
        base, out = rerun(n, transform=lambda x, d=nud: x * (1 + d))
        report(f"s (1 {nud:+g})", base, out)


def cmd_where(n):
    import collections
    fx = fixture()
    s, th = fx["s"], fx["theta"]
    sel = np.where(s * np.sin(th) < 0.5)[0][:n]
    L = O.lib()

    def one(i):
        d = O.make_dist(int(fx["kind"]), [float(p[i]) for p in fx["params"]])
        res = []
        for stokes in (1, 2):
            st = O.Stats()
            v = L.orc_heyvaerts(ctypes.byref(d), stokes, float(s[i]), float(th[i]), ctypes.byref(st))
            res.append((v, st.hey_fail_stage, st.hey_fail_outer_status, st.hey_fail_inner_status, st.hey_nan_kind,
                        st.hey_nan_x, st.hey_nan_mu, st.hey_fail_inner_var))
        return res

    with ThreadPoolExecutor(os.cpu_count()) as ex:
        out = list(ex.map(one, sel))
    cnt = collections.Counter()
    for i, r in zip(sel, out):
        for nm, (v, stage, ost, ist, nk, x, mu, iv) in zip(("rho_Q", "rho_V"), r):
            if np.isnan(v):
                cnt[(nm, f"stage {stage}", f"outer status {ost}", f"inner status {ist}",
                     "non-finite element: " + ("none" if nk == 0 else f"{'NR' if nk == 1 else 'QR'}, |mu| >= 1: {abs(mu) >= 1.0}, x < 1e-8: {abs(x) < 1e-8 or x != x}"))] += 1
    for k in sorted(cnt):
        print(cnt[k], *k)
    print("(stage 4 = QR march; outer status 5 = EFAILED: a NaN integrand value; 18 = EROUND; inner status 21 = ESING)")


def cmd_zone(idx):
    fx = fixture()
    L = O.lib()
    L.orc_test_hey_outer_integrand.restype = ctypes.c_double
    L.orc_test_hey_outer_integrand.argtypes = [ctypes.POINTER(O.Dist), ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                               ctypes.c_int, ctypes.c_double, ctypes.POINTER(O.Stats)]
    for i in idx:
        s, th = float(fx["s"][i]), float(fx["theta"][i])
        d = O.make_dist(int(fx["kind"]), [float(p[i]) for p in fx["params"]])
        print(f"point {i}: s = {s:.5g}, theta = {th:.4g}, p = {fx['params'][0][i]:.3g}, k = {fx['params'][1][i]:.3g}; "
              f"oracle rho_Q = {fx['out'][6][i]:.4g}; F(sigma) of rho_Q / live intervals of the inner QAG at sigma = s (1 +- 10^-j), j = 1..12")
        for sign in (-1, 1):
            row = []
            for j in range(1, 13):
                sg = s * (1 + sign * 10.0 ** (-j))
                if sg <= s * np.sin(th):
                    row.append("(outside)")
                    continue
                st = O.Stats()
                v = L.orc_test_hey_outer_integrand(ctypes.byref(d), 1, s, th, 1, sg, ctypes.byref(st))
                row.append(f"{v:.2e}/{st.max_gamma_intervals}")
            print("   " + ("+" if sign > 0 else "-"), " ".join(row))


def cmd_trace(i, stokes=1):
    fx = fixture()
    L = O.lib()
    L.orc_set_trace(1)
    d = O.make_dist(int(fx["kind"]), [float(p[i]) for p in fx["params"]])
    st = O.Stats()
    print(f"s = {fx['s'][i]!r}, theta = {fx['theta'][i]!r}, sigma0 = {fx['s'][i] * np.sin(fx['theta'][i])!r}", file=sys.stderr)
    v = L.orc_heyvaerts(ctypes.byref(d), stokes, float(fx["s"][i]), float(fx["theta"][i]), ctypes.byref(st))
    print(f"result {v}; failed in stage {st.hey_fail_stage}, outer status {st.hey_fail_outer_status}, inner status "
          f"{st.hey_fail_inner_status} at sigma = {st.hey_fail_inner_var:.17g}", file=sys.stderr)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "map"
    args = sys.argv[2:]
    if what == "map":
        cmd_map(args)
    elif what == "tolerance":
        cmd_tolerance(int(args[0]) if args else 400)
    elif what == "nudge":
        cmd_nudge(int(args[0]) if args else 640)
    elif what == "where":
        cmd_where(int(args[0]) if args else 300)
    elif what == "zone":
        cmd_zone([int(a) for a in args])
    elif what == "trace":
        cmd_trace(int(args[0]), int(args[1]) if len(args) > 1 else 1)
