"""Development check run on the GPU box: parity of the CUDA path against the
oracle fixtures + a first throughput number.  (The real tests live in tests/.)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rimphony_b200 as R  # noqa: E402

NAMES = R.COEFFICIENT_NAMES


def compare(tag, got, want, lobes):
    for c in range(8):
        a, b = got[c], want[c]
        both = np.isnan(a) & np.isnan(b)
        only_gpu = np.isnan(a) & ~np.isnan(b)
        only_orc = ~np.isnan(a) & np.isnan(b)
        ok = ~np.isnan(a) & ~np.isnan(b)
        if ok.sum() == 0:
            print(f"  {tag} {NAMES[c]:8s} no finite pairs (bothNaN {both.sum()})")
            continue
        rel = np.abs(a[ok] / b[ok] - 1)
        extra = ""
        if c in (4, 5):
            sc = np.abs(lobes[2 * (c - 4)]) + np.abs(lobes[2 * (c - 4) + 1])
            extra = " lobe-normalised max %.1e" % (np.abs(a[ok] - b[ok]) / sc[ok]).max()
        print(f"  {tag} {NAMES[c]:8s} median {np.median(rel):.1e} p99 {np.percentile(rel, 99):.1e} max {rel.max():.1e} "
              f">1e-3: {(rel > 1e-3).mean():.4f} | NaN both {both.sum()} gpu-only {only_gpu.sum()} oracle-only {only_orc.sum()} "
              f"sign-mismatch {(np.sign(a[ok]) != np.sign(b[ok])).sum()}{extra}")


def main():
    print("devices", R.device_count())
    print("fp64 peak TFLOP/s", R.fp64_peak_tflops())
    # 1. Bessel
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    n = 10 ** rng.uniform(np.log10(30), 10, 20000)
    eps = 10 ** rng.uniform(-12, 0, 20000)
    x = n * (1 - eps)
    n[:3000] = np.floor(rng.uniform(0, 30, 3000))
    x[:3000] = rng.uniform(0, 1, 3000) * n[:3000]
    j, dj = R.bessel_jn(n, x)
    rj = np.array([O.ref_bessel_j(a, b) for a, b in zip(n, x)])
    rdj = np.array([O.ref_bessel_dj(a, b) for a, b in zip(n, x)])
    sig = np.abs(rj) > 1e-30
    print("bessel J  rel err: median %.1e p99.9 %.1e max %.1e" % tuple(np.percentile(np.abs(j[sig] / rj[sig] - 1), [50, 99.9, 100])))
    sig = np.abs(rdj) > 1e-30
    print("bessel dJ rel err: median %.1e p99.9 %.1e max %.1e" % tuple(np.percentile(np.abs(dj[sig] / rdj[sig] - 1), [50, 99.9, 100])))
    print("small-int orders max abs err", np.abs(j[:3000] - rj[:3000]).max(), "max rel", np.nanmax(np.abs(j[:3000] / rj[:3000] - 1)[np.abs(rj[:3000]) > 1e-300]))

    # 2. parity against fixtures
    for name in ("pitchy_pl", "symphony_rows", "pitchy_kappa", "juettner_sweep"):
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        if not os.path.exists(path):
            print("fixture missing", name)
            continue
        fx = np.load(path)
        kind, s, th, params = int(fx["kind"]), fx["s"], fx["theta"], list(fx["params"])
        mask = 0xC0 if name == "juettner_sweep" else 0xFF
        for mode, tag in ((R.MODE_FAITHFUL, "faithful"), (R.MODE_FUSED, "fused")):
            t = time.time()
            res = R.compute_all_dimensionless_batch(kind, s, th, params, mode=mode, coeff_mask=mask, extras=True)
            dt = time.time() - t
            print(f"{name} [{tag}] n={len(s)} wall {dt:.2f}s kernels(ms) norm/sym/hey/total = {[round(v, 1) for v in res.kernel_ms]} "
                  f"status: nan {(res.status & 1).astype(bool).sum()} cap {(res.status & 2).astype(bool).sum()} "
                  f"GK apps mean sym {res.counters[0].mean():.0f} hey {res.counters[1].mean():.0f}")
            compare(tag, res.values, fx["out"], fx["lobes"])
            nrm = res.norm
            print("   norm finite:", np.isfinite(nrm).all())

    # 3. throughput
    for npts in (4096, 32768):
        kind, s, th, params = R.synthetic_batch("pitchy_pl", npts, seed=1)
        for mask, tag in ((0x3F, "symphony only"), (0xC0, "heyvaerts only"), (0xFF, "all 8")):
            t = time.time()
            res = R.compute_all_dimensionless_batch(kind, s, th, params, coeff_mask=mask, extras=True)
            dt = time.time() - t
            ms = res.kernel_ms
            print(f"throughput n={npts} {tag}: wall {dt:.2f}s -> {npts / dt:.0f} sets/s; kernel ms {[round(v, 1) for v in ms]}; "
                  f"GK apps/pt sym {res.counters[0].mean():.0f} hey {res.counters[1].mean():.0f}; "
                  f"status nan {(res.status & 1).astype(bool).mean():.3f} cap {(res.status & 2).astype(bool).mean():.4f}")
    print("launches", R.kernel_launch_count())


if __name__ == "__main__":
    main()
