#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
timeout 90 python tools/variant_bench.py 131072 pitchy_pl 2
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_nograde.so timeout 90 python tools/variant_bench.py 131072 pitchy_pl 2
} > $O/s11_variants.log 2>&1
cat $O/s11_variants.log
