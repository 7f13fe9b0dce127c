#!/bin/bash
# GPU session: software-cohort lock-step variants A/B (short timeouts: a hang must not eat the budget)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
for v in s16x8 s16x4 s16x16 s20x10 s20x5; do
  RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_$v.so timeout 75 python tools/variant_bench.py 131072 pitchy_pl 2 || echo "variant $v failed rc=$?"
done
} > $O/s4_variants.log 2>&1
v=s16x8
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_$v.so timeout 240 ncu --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 -o $O/s4_${v}_warpstate -f python tools/profile_small.py 8192 0xFF > $O/s4_ncu_$v.log 2>&1
timeout 300 python -m pytest tests/test_crank_out.py tests/test_examples.py -m gpu -x -q > $O/s4_pytest_small.log 2>&1
cat $O/s4_variants.log
