#!/bin/bash
# lean math (rb_div / rb_exp / rb_log, tables out of the literal pool) against the head build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_base.so timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2
} > $O/s12_variants.log 2>&1
timeout 600 python tools/fast_check.py 0xFF 32768 > $O/s12_fast_check.log 2>&1
cat $O/s12_variants.log; grep -A8 "pitchy_pl_10k\|^throughput" $O/s12_fast_check.log | cut -c1-260
