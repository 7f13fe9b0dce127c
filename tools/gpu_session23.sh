#!/bin/bash
# Symphony kernel without its outer tile: CTAs per SM; plus the relaxed I-series truncation and shared square roots
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2
for v in s6 s8 s10; do RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_$v.so timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2; done
} > $O/s23_variants.log 2>&1
grep -E " (sym|hey|all):" $O/s23_variants.log
