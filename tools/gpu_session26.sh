#!/bin/bash
# Heyvaerts: CTAs per SM against DRAM traffic (local memory beyond the L2 at 10 CTAs x 768 B of stack per thread)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for v in default h8 h9; do
  L=""; [ $v != default ] && L=$PWD/rimphony_b200/variants/librimphony_b200_$v.so
  RIMPHONY_B200_LIB=$L timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2 > $O/s26_$v.log 2>&1
  RIMPHONY_B200_LIB=$L timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum --clock-control none -k regex:'k_heyvaerts_fast' -c 1 --csv --log-file $O/s26_dram_$v.csv python tools/profile_small.py 65536 0xC0 > $O/s26_dram_${v}_run.log 2>&1
  grep -E " (hey|all):" $O/s26_$v.log; grep -E "dram__bytes|gpu__time" $O/s26_dram_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
