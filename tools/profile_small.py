"""Tiny driver for ncu: one batched call on a small seeded pitchy power-law batch.
usage: python tools/profile_small.py N_POINTS COEFF_MASK [MODE]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rimphony_b200 as R  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1776
mask = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0xFF
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
kind, s, theta, params = R.synthetic_batch("pitchy_pl", n, seed=1)
res = R.compute_all_dimensionless_batch(kind, s, theta, params, coeff_mask=mask, mode=mode, extras=True)
print("n", n, "mask", hex(mask), "kernel ms", [round(v, 2) for v in res.kernel_ms], "apps/pt",
      res.counters[0].mean(), res.counters[1].mean())
