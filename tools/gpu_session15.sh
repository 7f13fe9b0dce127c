#!/bin/bash
# which of the second batch of changes made the Heyvaerts kernel slower (632 -> 670 ms)?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
for v in fwd ieeesqrt both cterms; do RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_$v.so timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2; done
} > $O/s15_variants.log 2>&1
grep -E " (hey|sym):" $O/s15_variants.log
