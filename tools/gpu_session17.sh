#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_b4.so timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2
} > $O/s17_variants.log 2>&1
cat $O/s17_variants.log
