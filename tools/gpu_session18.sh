#!/bin/bash
# Heyvaerts kernel: CTAs per SM now that its tiles are a quarter of the size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2
for v in h6 h7 h8 h10; do RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_$v.so timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2; done
} > $O/s18_variants.log 2>&1
grep -E " (hey|all):" $O/s18_variants.log
