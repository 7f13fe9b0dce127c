#!/bin/bash
# ncu --set full with source of the lean-math build (where the instructions go now)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 -o $O/s13_fast_full -f python tools/profile_small.py 8192 0xFF > $O/s13_ncu_full.log 2>&1
tail -3 $O/s13_ncu_full.log
