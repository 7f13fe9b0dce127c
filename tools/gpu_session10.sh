#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
timeout 90 python tools/variant_bench.py 131072 pitchy_pl 2
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_prev.so timeout 90 python tools/variant_bench.py 131072 pitchy_pl 2
} > $O/s10_variants.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/s10_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s10_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/s10_bench_c3.json 2> $O/s10_bench_c3.err
cat $O/s10_variants.log; tail -8 $O/s10_pytest.log | cut -c1-300; python -c "
import json; d=json.load(open('$O/s10_bench_c3.json')); print(round(d['value']), round(d['e2e']['value']), d['roofline']['frac'], d['roofline']['gk31_applications_per_point'], d['parity']['meets_north_star_but_for_reference_failures'])"
