"""How well defined is the reference's own output?  (TEST INFRASTRUCTURE: runs the CPU oracle.)

    python tests/golden/make_stability.py FIXTURE [MASK]

For every point of a golden fixture the oracle (the reference's algorithm, oracle/) is re-run twice:

  tight   every QAG call of the path with epsrel 3e-4 instead of the reference's 1e-3
          (oracle.set_epsrel); a converged adaptive integral moves by less than its tolerance
          when the tolerance is tightened;
  nudge   the same algorithm at s (1 + 1e-9): the exact coefficients move by ~1e-9.

Where the three runs (fixture, tight, nudge) disagree by more than 1e-3, or one of them is NaN
(a QAG failure) and another is not, the reference's value is set by where its nested adaptive
quadrature happens to put its nodes, not by the integral: parity "to 1e-3 of the reference" is not
defined for that coefficient of that point.  tests/ and bench.py use the mask
`reference_defined` built from these arrays (rimphony_b200.parity.reference_defined) and report
the masked fraction next to every parity figure.

Writes tests/golden/<FIXTURE>_stability.npz: tight[8, n], nudge[8, n] (NaN where not requested).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as O  # noqa: E402

TIGHT = 3e-4
NUDGE = 1e-9


def main():
    name = sys.argv[1]
    mask = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0xFF
    fx = np.load(os.path.join(HERE, name + ".npz"))
    kind, s, theta, params = int(fx["kind"]), fx["s"], fx["theta"], list(fx["params"])
    t0 = time.time()
    O.set_epsrel(TIGHT, TIGHT)
    tight, tight_lobes = O.batch(kind, s, theta, params, coeff_mask=mask)
    O.set_epsrel(0.0, 0.0)
    t1 = time.time()
    nudge, nudge_lobes = O.batch(kind, s * (1.0 + NUDGE), theta, params, coeff_mask=mask)
    t2 = time.time()
    path = os.path.join(HERE, name + "_stability.npz")
    keep = dict(np.load(path)) if os.path.exists(path) else {}   # e.g. the `converged` arrays of make_converged.py
    keep.update(tight=tight, nudge=nudge, tight_lobes=tight_lobes, nudge_lobes=nudge_lobes, tight_epsrel=TIGHT,
                nudge_rel=NUDGE, mask=mask)
    np.savez_compressed(path, **keep)
    print(f"{name}: tight {t1 - t0:.0f} s, nudge {t2 - t1:.0f} s")
    base = fx["out"]
    for c in range(8):
        if not (mask >> c) & 1:
            continue
        a, b, d = base[c], tight[c], nudge[c]
        fin = np.isfinite(a) & np.isfinite(b) & np.isfinite(d)
        anynan = ~(np.isfinite(a) & np.isfinite(b) & np.isfinite(d))
        allnan = ~np.isfinite(a) & ~np.isfinite(b) & ~np.isfinite(d)
        sc = np.abs(a)
        spread = np.maximum(np.abs(a - b), np.abs(a - d)) / sc
        print(f"  slot {c}: all finite {fin.sum()}, all NaN {allnan.sum()}, NaN in some runs only {(anynan & ~allnan).sum()}, "
              f"finite but spread > 1e-3: {(fin & (spread > 1e-3)).sum()}  (> 1e-2: {(fin & (spread > 1e-2)).sum()})")


if __name__ == "__main__":
    main()
