"""Development check run on the GPU box: the product (FAST) path against the oracle
fixtures, per coefficient, plus kernel timings.  (The real tests live in tests/.)
usage: python tools/fast_check.py [MASK] [N_THROUGHPUT]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rimphony_b200 as R  # noqa: E402

NAMES = R.COEFFICIENT_NAMES


def compare(got, want, lobes, mask):
    for c in range(8):
        if not (mask >> c) & 1:
            continue
        a, b = got[c], want[c]
        both = np.isnan(a) & np.isnan(b)
        only_gpu = np.isnan(a) & ~np.isnan(b)
        only_orc = ~np.isnan(a) & np.isnan(b)
        ok = ~np.isnan(a) & ~np.isnan(b)
        if ok.sum() == 0:
            print(f"  {NAMES[c]:8s} no finite pairs (bothNaN {both.sum()})")
            continue
        rel = np.abs(a[ok] / b[ok] - 1)
        extra = ""
        if c in (4, 5):
            sc = np.abs(lobes[2 * (c - 4)]) + np.abs(lobes[2 * (c - 4) + 1])
            reln = np.abs(a[ok] - b[ok]) / sc[ok]
            extra = " | lobe-normalised p99.9 %.1e max %.1e >1e-3: %.5f" % (np.percentile(reln, 99.9), reln.max(), (reln > 1e-3).mean())
        print(f"  {NAMES[c]:8s} median {np.median(rel):.1e} p99 {np.percentile(rel, 99):.1e} p99.9 {np.percentile(rel, 99.9):.1e} "
              f"max {rel.max():.1e} >1e-3: {(rel > 1e-3).mean():.5f} | NaN both {both.sum()} gpu-only {only_gpu.sum()} "
              f"oracle-only {only_orc.sum()} sign-mismatch {(np.sign(a[ok]) != np.sign(b[ok])).sum()}{extra}")


def main():
    mask = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0xFF
    n_tp = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    only_tp = os.environ.get("FAST_CHECK_ONLY_THROUGHPUT") == "1"
    for name in (() if only_tp else ("pitchy_pl", "symphony_rows", "powerlaw", "pitchy_kappa", "juettner_sweep", "pitchy_pl_4k", "pitchy_pl_10k", "pitchy_kappa_10k", "powerlaw_10k")):
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        if not os.path.exists(path):
            continue
        fx = np.load(path)
        kind, s, th, params = int(fx["kind"]), fx["s"], fx["theta"], list(fx["params"])
        m = (0xC0 if name == "juettner_sweep" else 0xFF) & mask
        if not m:
            continue
        t = time.time()
        res = R.compute_all_dimensionless_batch(kind, s, th, params, mode=R.MODE_FAST, coeff_mask=m, extras=True)
        dt = time.time() - t
        print(f"{name} [fast] n={len(s)} wall {dt:.2f}s kernels(ms) norm/sym/hey/total = {[round(v, 1) for v in res.kernel_ms]} "
              f"status: nan {(res.status & 1).astype(bool).sum()} cap {(res.status & 2).astype(bool).sum()} "
              f"GK apps mean sym {res.counters[0].mean():.0f} hey {res.counters[1].mean():.0f}")
        compare(res.values, fx["out"], fx["lobes"], m)

    for cfg in (("pitchy_pl",) if only_tp else ("pitchy_pl", "pitchy_kappa")):
        n_cfg = n_tp if cfg == "pitchy_pl" else max(1024, n_tp // 8)  # kappa: ~23 % of points take the faithful route
        kind, s, th, params = R.synthetic_batch(cfg, n_cfg, seed=1)
        for mk, tag in ((0x3F, "symphony only"), (0xC0, "heyvaerts only"), (0xFF, "all 8")):
            if not (mk & mask) or (mk == 0xFF and mask != 0xFF):
                continue
            t = time.time()
            res = R.compute_all_dimensionless_batch(kind, s, th, params, coeff_mask=mk, extras=True)
            dt = time.time() - t
            ms = res.kernel_ms
            print(f"throughput {cfg} n={n_cfg} {tag}: wall {dt:.2f}s -> {n_cfg / dt:.0f} sets/s; kernel ms {[round(v, 1) for v in ms]}; "
                  f"GK apps/pt sym {res.counters[0].mean():.0f} hey {res.counters[1].mean():.0f}; "
                  f"status nan {(res.status & 1).astype(bool).mean():.3f} cap {(res.status & 2).astype(bool).mean():.4f}")
    print("launches", R.kernel_launch_count())


if __name__ == "__main__":
    main()
