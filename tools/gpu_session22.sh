#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 -o $O/s22_fast_full -f python tools/profile_small.py 16384 0xFF > $O/s22_ncu_full.log 2>&1
tail -2 $O/s22_ncu_full.log
