#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/s9_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s9_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/s9_smoke.log 2>&1; echo "smoke rc=$?" >> $O/s9_smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/s9_bench_ref.json 2> $O/s9_bench_ref.err
tail -15 $O/s9_pytest.log | cut -c1-300; cat $O/s9_smoke.log; cat $O/s9_bench_ref.json | cut -c1-900
