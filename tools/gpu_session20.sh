#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2 > $O/s20_variants.log 2>&1
timeout 120 python tools/variant_bench.py 131072 powerlaw 1 >> $O/s20_variants.log 2>&1
cat $O/s20_variants.log
