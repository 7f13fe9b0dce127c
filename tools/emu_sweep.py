"""Development loop without a GPU: the product (FAST) path's device headers, compiled for the host
(tests/hostemu, lanes as loops), run over a golden fixture on all host cores and compared with the
fixture's oracle outputs.  Not a product path, not a fallback: nothing in the package can load it.
usage: python tools/emu_sweep.py FIXTURE [which=3] [mode=2] [first_n]"""
import ctypes
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["j_I", "alpha_I", "j_Q", "alpha_Q", "j_V", "alpha_V", "rho_Q", "rho_V"]


def load_emu():
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "hostemu")], check=True)
    lib = ctypes.CDLL(os.environ.get("EMU_LIB") or os.path.join(ROOT, "tests", "hostemu", "_build", "libhostemu.so"))
    dp, up = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint)
    lib.emu_point.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                              dp, dp, dp, up]
    return lib


def run_fixture(lib, fx, which=3, mode=2, sel=None, eps=(1e-3, 1e-3, 1e-3, 1e-3), threads=None):
    kind = int(fx["kind"])
    s, th, params = fx["s"], fx["theta"], fx["params"]
    idx = np.arange(len(s)) if sel is None else np.asarray(sel)
    out = np.full((8, len(idx)), np.nan)
    lobes = np.full((4, len(idx)), np.nan)
    info = np.zeros((4, len(idx)), dtype=np.uint32)

    def one(k):
        i = idx[k]
        pv = (ctypes.c_double * params.shape[0])(*[float(p[i]) for p in params])
        ev = (ctypes.c_double * 4)(*eps)
        o, l, nf = (ctypes.c_double * 8)(), (ctypes.c_double * 4)(), (ctypes.c_uint * 4)()
        lib.emu_point(kind, pv, params.shape[0], mode, which, float(s[i]), float(th[i]), ev, o, l, nf)
        out[:, k] = o[:]
        lobes[:, k] = l[:]
        info[:, k] = nf[:]

    with ThreadPoolExecutor(threads or os.cpu_count()) as ex:
        list(ex.map(one, range(len(idx))))
    return out, lobes, info


def compare(got, want, lobes, mask=0xFF, label=""):
    for c in range(8):
        if not (mask >> c) & 1:
            continue
        a, b = got[c], want[c]
        both = np.isnan(a) & np.isnan(b)
        only_a = np.isnan(a) & ~np.isnan(b)
        only_b = ~np.isnan(a) & np.isnan(b)
        ok = ~np.isnan(a) & ~np.isnan(b)
        if ok.sum() == 0:
            print(f"  {label}{NAMES[c]:8s} no finite pairs (bothNaN {both.sum()} here-only {only_a.sum()} oracle-only {only_b.sum()})")
            continue
        sc = np.abs(b)
        if c in (4, 5):
            sc = np.abs(lobes[2 * (c - 4)]) + np.abs(lobes[2 * (c - 4) + 1])
        rel = np.abs(a[ok] - b[ok]) / sc[ok]
        print(f"  {label}{NAMES[c]:8s} n {len(a)} median {np.median(rel):.1e} p99 {np.percentile(rel, 99):.1e} max {rel.max():.1e} "
              f">1e-3: {(rel > 1e-3).sum()} ({(rel > 1e-3).mean():.5f}) | NaN both {both.sum()} here-only {only_a.sum()} "
              f"oracle-only {only_b.sum()} sign {(np.sign(a[ok]) != np.sign(b[ok])).sum()}")


def main():
    name = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    first = int(sys.argv[4]) if len(sys.argv) > 4 else None
    fx = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    lib = load_emu()
    sel = None if first is None else np.arange(first)
    t = time.time()
    out, lobes, info = run_fixture(lib, fx, which, mode, sel)
    dt = time.time() - t
    n = out.shape[1]
    print(f"{name}: {n} points in {dt:.1f} s; applications/pt sym {info[0].mean():.0f} hey {info[2].mean():.0f}")
    want = fx["out"][:, :n] if sel is None else fx["out"][:, sel]
    wl = fx["lobes"][:, :n] if sel is None else fx["lobes"][:, sel]
    compare(out, want, wl, (0x3F if which & 1 else 0) | (0xC0 if which & 2 else 0))
    if which & 2:
        sigma0 = (fx["s"] * np.sin(fx["theta"]))[:n]
        for lo, hi in ((0, 0.5), (0.5, 1), (1, 3), (3, 10), (10, 1e9)):
            m = (sigma0 >= lo) & (sigma0 < hi)
            if m.sum():
                print(f" sigma0 in [{lo}, {hi}): {m.sum()} points, applications/pt {info[2][m].mean():.0f}")
                compare(out[:, m], want[:, m], wl[:, m], 0xC0, "  ")
    np.savez("/tmp/emu_" + name + ".npz", out=out, lobes=lobes, info=info)


if __name__ == "__main__":
    main()
