#!/bin/bash
# lean math in the faithful Symphony kernels (the pitchy-kappa continuation): parity and time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
timeout 300 python tools/kappa_check.py
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_symlean.so timeout 300 python tools/kappa_check.py
} > $O/s24_kappa.log 2>&1
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_symlean.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/s24_pytest_symlean.log 2>&1; echo "rc=$?" >> $O/s24_pytest_symlean.log
cat $O/s24_kappa.log; tail -5 $O/s24_pytest_symlean.log | cut -c1-300
