#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
for v in h10 h12 h16; do RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_$v.so timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2; done
RIMPHONY_B200_SERIAL_STAGES=1 RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_h10.so timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2
} > $O/s19_variants.log 2>&1
grep -E " (sym|hey|all):" $O/s19_variants.log
