"""The batched crank-out tool writes the reference's on-disk format (SURVEY section 8 row f1;
examples/crank-out-pitchypl.rs:132-155, 175-194; crank-out-pitchykappa.rs:165-216)."""
import math
import re

import numpy as np
import pytest

from rimphony_b200 import crank_out

ROW = re.compile(r"^(-?\d\.\d{16}e-?\d+|NaN|inf|-inf)(\t(-?\d\.\d{16}e-?\d+|NaN|inf|-inf))*$")


def test_rust_scientific_format():
    assert crank_out.rust_sci(1.0) == "1.0000000000000000e0"
    assert crank_out.rust_sci(-0.15625) == "-1.5625000000000000e-1"
    assert crank_out.rust_sci(1e10) == "1.0000000000000000e10"
    assert crank_out.rust_sci(float("nan")) == "NaN"
    assert float(crank_out.rust_sci(math.pi)) == math.pi  # 17 significant digits round-trip


def fake_eval(tool, cols, n_devices):
    n = len(cols["s"])
    vals = np.outer(np.arange(1, 9), cols["s"])
    vals[6, ::3] = np.nan
    return vals


@pytest.mark.parametrize("tool,argv,header", [
    ("pitchypl", ["0.07", "1e4", "0.003", "1.5705", "1.5", "4", "0", "3"], crank_out.PITCHYPL_HEADER),
    ("pitchykappa", ["1", "1e6", "0.003", "1.5705", "1.5", "4.5", "2.718", "20.08", "0", "3"], crank_out.PITCHYKAPPA_HEADER),
])
def test_tsv_layout_and_append(tmp_path, tool, argv, header):
    path = tmp_path / "out.txt"
    for _ in range(2):  # the second run appends and re-emits the header, like the reference
        crank_out.main(["--block", "7", "--count", "17", "--seed", "3", tool] + argv + [str(path)], evaluate_fn=fake_eval)
    lines = path.read_text().splitlines()
    assert len(lines) == 2 * (1 + 17)
    assert lines[0] == "\t".join(header) and lines[18] == lines[0]
    n_par = len(header) - 9
    for ln in lines[1:18]:
        assert ROW.match(ln), ln
        f = ln.split("\t")
        assert len(f) == len(header)
        s = float(f[0])
        lo, hi = float(argv[0]), float(argv[1])
        assert lo <= s <= hi
        assert float(f[n_par + 1 + 2]) == pytest.approx(3 * s, rel=1e-15)  # j_Q column of the fake evaluator
    assert sum("NaN" in ln for ln in lines[1:18]) >= 5


def test_demo_powerlaw_layout():
    """examples/demo-powerlaw.rs:63-99, 123-163."""
    import io
    buf = io.StringIO()
    crank_out.run_demo("almostuniform1", buf, evaluate_fn=lambda cols: np.outer(np.arange(1, 9), cols["s"]))
    lines = buf.getvalue().splitlines()
    assert lines[0] == "\t".join(crank_out.DEMO_HEADER) and len(lines) == 65
    first, last = [float(v) for v in lines[1].split("\t")], [float(v) for v in lines[-1].split("\t")]
    assert first[:6] == [100.0, 0.5, 3.0, 0.0, 0.0, 1e5]
    assert last[:6] == pytest.approx([90.0, 0.6, 2.5, 3e10, 0.1, 7e4], rel=1e-15)
    assert all(ROW.match(ln) for ln in lines[1:])


def _against_oracle(rows8, fx):
    """rows8: [n, 8] from the tool; fx: an oracle fixture of the same points.  Bars of the FAST mode."""
    want, lobes = fx["out"], fx["lobes"]
    for c in range(8):
        a, b = rows8[:, c], want[c]
        assert np.array_equal(np.isnan(a), np.isnan(b)), c
        scale = np.abs(b)
        if c in (4, 5):
            scale = np.abs(lobes[2 * (c - 4)]) + np.abs(lobes[2 * (c - 4) + 1])
        err = np.abs(a - b) / scale
        # the reference's own nested-QAG noise reaches ~1e-3: at most one row beyond it, none beyond 3e-3
        assert (err > 1e-3).sum() <= 1 and np.nanmax(err) <= 3e-3, (c, np.nanmax(err))
        assert (np.sign(a) == np.sign(b))[np.abs(b) > 1e-2 * scale].all()


@pytest.mark.gpu
def test_demo_powerlaw_on_the_device(golden):
    """examples/demo-powerlaw.rs on the GPU: the 64 rows against the oracle's (fixture demo_powerlaw)."""
    import io
    buf = io.StringIO()
    crank_out.run_demo("almostuniform1", buf)
    rows = np.array([[float(x) for x in ln.split("\t")] for ln in buf.getvalue().splitlines()[1:]])
    fx = golden("demo_powerlaw")
    assert rows.shape == (64, 15)
    assert np.array_equal(rows[:, 0], fx["s"]) and np.array_equal(rows[:, 1], fx["theta"]) and np.array_equal(rows[:, 2], fx["params"][0])
    _against_oracle(rows[:, 7:], fx)


@pytest.mark.gpu
def test_crank_out_on_the_device(tmp_path, golden):
    """crank-out-pitchypl on the GPU: the rows a seeded run writes against the oracle's values for the same
    points (fixture crank_pitchypl_64, drawn by the same sampler calls)."""
    path = tmp_path / "pl.txt"
    crank_out.main(["--block", "64", "--count", "64", "--seed", "1", "pitchypl", "1", "100", "0.3", "1.5", "2", "3", "0", "2",
                    str(path)])
    rows = np.array([[float(x) for x in ln.split("\t")] for ln in path.read_text().splitlines()[1:]])
    assert rows.shape == (64, 13)
    fx = golden("crank_pitchypl_64")
    for j, col in enumerate((fx["s"], fx["theta"], fx["params"][0], fx["params"][1])):
        assert np.array_equal(rows[:, j], col)   # {:.16e} round-trips a double
    _against_oracle(rows[:, 5:], fx)
