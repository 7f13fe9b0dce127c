#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/s6_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s6_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/s6_bench_c3.json 2> $O/s6_bench_c3.err
timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2 > $O/s6_kernels.log 2>&1
timeout 300 python bench.py --config powerlaw --points 262144 --steps 2 --warmup 1 --no-cpu-baseline > $O/s6_bench_c2.json 2> $O/s6_bench_c2.err
tail -40 $O/s6_pytest.log | cut -c1-300; cat $O/s6_kernels.log; cat $O/s6_bench_c3.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d['roofline']['frac']); print(json.dumps(d['parity'])[:3000])"
