#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
timeout 90 python tools/variant_bench.py 131072 pitchy_pl 2
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_smemctx.so timeout 90 python tools/variant_bench.py 131072 pitchy_pl 2
} > $O/s7_variants.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_smemctx.so timeout 200 ncu --metrics $M --clock-control none -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 --csv --log-file $O/s7_smemctx.csv python tools/profile_small.py 8192 0xFF > $O/s7_smemctx_run.log 2>&1
cat $O/s7_variants.log; grep -E "dram__bytes|hit_rate" $O/s7_smemctx.csv | cut -d, -f5,13-16
