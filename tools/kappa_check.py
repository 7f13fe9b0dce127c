"""Development check on the GPU box: the pitchy-kappa 1e4-point fixture (C4) with the library selected by
RIMPHONY_B200_LIB: parity of j / alpha, share of points that took the faithful continuation, kernel time."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rimphony_b200 as R  # noqa: E402
from rimphony_b200 import parity as P  # noqa: E402

fx = P.load_fixture("pitchy_kappa_10k")
res = R.compute_all_dimensionless_batch(int(fx["kind"]), fx["s"], fx["theta"], list(fx["params"]), coeff_mask=0x3F, extras=True)
st = P.parity_stats(res.values, fx["out"], fx["lobes"], fx["defined"], mask=0x3F)
tag = os.path.basename(os.environ.get("RIMPHONY_B200_LIB", "default"))
print(f"{tag}: kernel ms {[round(v, 1) for v in res.kernel_ms]} -> {len(fx['s']) / res.kernel_ms[3] * 1e3:.0f} sets/s (Symphony only); "
      f"rerouted {((res.status & 8) != 0).mean():.4f}; apps/pt {res.counters[0].mean():.0f}")
for k, v in st.items():
    print(f"   {k:8s} within {v['frac_within']:.5f} max {v['max_err']:.2e} nan here-only {v['nan_here_only']} ref-only {v['nan_ref_only']}")
