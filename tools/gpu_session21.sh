#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/s21_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s21_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/s21_bench_c3.json 2> $O/s21_bench_c3.err
tail -8 $O/s21_pytest.log | cut -c1-300; python -c "
import json; d=json.load(open('$O/s21_bench_c3.json')); print(round(d['value']), round(d['e2e']['value']), d['roofline'], d['parity'])"
