#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_onetest.so timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2
} > $O/s16_variants.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 -o $O/s16_fast_full -f python tools/profile_small.py 8192 0xFF > $O/s16_ncu_full.log 2>&1
cat $O/s16_variants.log; tail -2 $O/s16_ncu_full.log
