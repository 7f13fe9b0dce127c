#!/bin/bash
# lean sqrt, backward I-series, paired Miller recurrence, Debye pair: kernel times and parity of the 10k fixtures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2 > $O/s14_variants.log 2>&1
timeout 600 python tools/fast_check.py 0xFF 32768 > $O/s14_fast_check.log 2>&1
cat $O/s14_variants.log; grep -A8 "pitchy_pl_10k" $O/s14_fast_check.log | cut -c1-260
