"""Parity statistics against golden fixtures (pure numpy; reads committed .npz files only).

The north-star bar is "<= 1e-3 relative on >= 99.9 % of points, NaN where the reference is NaN,
same signs" against the reference's CPU/GSL algorithm.  The fixtures under tests/golden/ hold that
algorithm's outputs (the oracle) for seeded inputs; the `*_stability.npz` companions hold two more
runs of the SAME algorithm (every QAG tolerance 3e-4 instead of 1e-3; s nudged by 1e-9), made by
tests/golden/make_stability.py.  A coefficient of a point is *reference-defined* when the three
runs agree: all finite and within `tol` of each other, or all NaN.  Where they do not (a QAG
failure that comes and goes, a value that moves by more than 1e-3 when the reference's own
tolerance is tightened) the reference's output is an artefact of where its nested adaptive
quadrature happens to put its nodes, and "parity to 1e-3" has no meaning; those entries are
counted and reported (`undefined`), never silently dropped.  A third kind of evidence
(tests/golden/make_converged.py): for the entries where the product path's algorithm and the
fixture disagree, the same reference algorithm at epsrel 1e-5; where THAT differs from the
fixture by more than 1e-3 the reference's own integration error exceeds its nominal tolerance
(QUADPACK's error estimate is a heuristic), and the entry is reference-undefined as well.

Nothing here imports the oracle: bench.py and the tests call this with arrays they loaded.
"""
import os

import numpy as np

COEFFICIENT_NAMES = ("j_I", "alpha_I", "j_Q", "alpha_Q", "j_V", "alpha_V", "rho_Q", "rho_V")
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def scales(out, lobes):
    """Per-entry error scale: |value|, except Stokes V where the two lobes that the reference
    integrates separately nearly cancel (symphony.rs:97-107): |lobe+| + |lobe-|."""
    sc = np.abs(np.asarray(out, dtype=np.float64)).copy()
    if lobes is not None:
        sc[4] = np.abs(lobes[0]) + np.abs(lobes[1])
        sc[5] = np.abs(lobes[2]) + np.abs(lobes[3])
    return sc


def reference_defined(base, lobes=None, variants=(), tol=1e-3, mask=0xFF, converged=None, converged_set=None):
    """[8, n] bool: True where every run of the reference's algorithm agrees (see module doc).
    `mask`: the slots the variants were computed for (the others count as defined).
    `converged` / `converged_set`: the same algorithm at epsrel 1e-5 for the entries in `converged_set`
    (tests/golden/make_converged.py): where the fixture value is further than `tol` from its own converged
    limit, the reference's integration error exceeds what is asked of others."""
    base = np.asarray(base, dtype=np.float64)
    ok = np.ones(base.shape, dtype=bool)
    sc = scales(base, lobes)
    if converged is not None and converged_set is not None:
        with np.errstate(invalid="ignore", divide="ignore"):
            far = np.abs(base - np.asarray(converged, dtype=np.float64)) > tol * sc   # a failed (NaN) run is no evidence
        ok &= ~(np.asarray(converged_set, dtype=bool) & far)
    for v in variants:
        v = np.asarray(v, dtype=np.float64)
        for c in range(8):
            if not (mask >> c) & 1:
                continue
            a, b = base[c], v[c]
            both_nan = np.isnan(a) & np.isnan(b)
            both_fin = np.isfinite(a) & np.isfinite(b)
            with np.errstate(invalid="ignore", divide="ignore"):
                close = both_fin & (np.abs(a - b) <= tol * sc[c])
            ok[c] &= both_nan | close
    return ok


def load_fixture(name):
    """A golden fixture and, when present, its stability companion."""
    fx = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    st_path = os.path.join(GOLDEN_DIR, name + "_stability.npz")
    variants = ()
    st_mask = 0
    conv = cset = None
    if os.path.exists(st_path):
        st = np.load(st_path)
        variants = (st["tight"], st["nudge"])
        st_mask = int(st["mask"])
        if "converged" in st:
            conv, cset = st["converged"], st["converged_set"]
    fx["defined"] = reference_defined(fx["out"], fx.get("lobes"), variants, mask=st_mask, converged=conv, converged_set=cset)
    fx["stability_mask"] = st_mask
    fx["converged_entries"] = int(cset.sum()) if cset is not None else 0
    return fx


def parity_stats(got, want, lobes=None, defined=None, mask=0xFF, tol=1e-3):
    """Per-slot statistics of `got` against the fixture outputs `want` ([8, n] each).

    Returns {slot name: {n, undefined, nan_both, nan_mismatch, nan_here_only, nan_ref_only, finite,
    within, frac_within, max_err, sign_mismatch}}: `undefined` entries (see reference_defined) are excluded from
    the other counts; `frac_within` is over the finite, defined pairs; `nan_mismatch` counts
    defined entries that are NaN on exactly one side."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    sc = scales(want, lobes)
    n = want.shape[1]
    if defined is None:
        defined = np.ones(want.shape, dtype=bool)
    res = {}
    for c in range(8):
        if not (mask >> c) & 1:
            continue
        d = defined[c]
        a, b = got[c], want[c]
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        fin = d & ~nan_a & ~nan_b
        with np.errstate(invalid="ignore", divide="ignore"):
            err = np.abs(a - b) / sc[c]
        err_f = err[fin]
        within = int((err_f <= tol).sum())
        res[COEFFICIENT_NAMES[c]] = {
            "n": int(n),
            "undefined": int((~d).sum()),
            "nan_both": int((d & nan_a & nan_b).sum()),
            "nan_mismatch": int((d & (nan_a != nan_b)).sum()),
            "nan_here_only": int((d & nan_a & ~nan_b).sum()),    # a failure here where the reference has a number
            "nan_ref_only": int((d & ~nan_a & nan_b).sum()),     # the reference's QAG gave up, this path did not
            "finite": int(fin.sum()),
            "within": within,
            "frac_within": (within / int(fin.sum())) if fin.any() else None,
            "max_err": float(err_f.max()) if fin.any() else None,
            "sign_mismatch": int((np.sign(a[fin]) != np.sign(b[fin])).sum()),
        }
    return res


def summarize(stats):
    """The compact `parity` object of bench.py's JSON line."""
    names = list(stats)
    first = stats[names[0]]
    return {
        "n": first["n"],
        "slots": names,
        "within_1e-3": [None if stats[k]["frac_within"] is None else round(stats[k]["frac_within"], 5) for k in names],
        "nan_mismatch": [stats[k]["nan_mismatch"] for k in names],
        "nan_here_only": [stats[k]["nan_here_only"] for k in names],
        "nan_reference_only": [stats[k]["nan_ref_only"] for k in names],
        "sign_mismatch": [stats[k]["sign_mismatch"] for k in names],
        "max_err": [None if stats[k]["max_err"] is None else float(f"{stats[k]['max_err']:.3g}") for k in names],
        "reference_undefined": [stats[k]["undefined"] for k in names],
    }


def meets_north_star(stats, frac=0.999, nan_frac=0.001, reference_failures_allowed=0.0):
    """True when every slot is within 1e-3 on >= 99.9 % of its finite defined pairs, NaN
    patterns differ on <= 0.1 % of the points and no sign differs.  With
    `reference_failures_allowed` (a fraction of the points) entries where the REFERENCE returned
    its failure marker and this path a number are tolerated up to that fraction; a NaN here
    where the reference has a number is never tolerated beyond `nan_frac`."""
    for k, v in stats.items():
        if v["frac_within"] is not None and v["frac_within"] < frac:
            return False
        if v["sign_mismatch"] or v["nan_here_only"] > nan_frac * v["n"]:
            return False
        if v["nan_ref_only"] > max(nan_frac, reference_failures_allowed) * v["n"]:
            return False
    return True
