#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
for v in ls16 ls16ns ls12ns ls8x2ns; do
  RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_$v.so timeout 75 python tools/variant_bench.py 131072 pitchy_pl 2 || echo "variant $v failed rc=$?"
done
} > $O/s5_variants.log 2>&1
cat $O/s5_variants.log
