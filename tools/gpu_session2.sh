#!/bin/bash
# GPU session: test-suite, bench lines of every BASELINE config, ncu captures of the head build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/s2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s2_pytest.log
python bench.py --steps 5 --warmup 3 > $O/s2_bench_c3.json 2> $O/s2_bench_c3.err
python bench.py --config powerlaw --points 1000000 --steps 2 --warmup 1 > $O/s2_bench_c2.json 2> $O/s2_bench_c2.err
python bench.py --config juettner_sweep --steps 5 --warmup 3 > $O/s2_bench_c5.json 2> $O/s2_bench_c5.err
python bench.py --config pitchy_kappa --points 16384 --steps 1 --warmup 1 > $O/s2_bench_c4.json 2> $O/s2_bench_c4.err
python bench.py --single-process --gpus 1 --points 262144 --steps 2 --warmup 1 --no-parity > $O/s2_bench_sp1.json 2> $O/s2_bench_sp1.err
# ncu: counters for the roofline constants (after the plain run above exited 0), then the full sets
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
python tools/profile_small.py 8192 0xFF > $O/s2_consts_plain.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 --csv --log-file $O/s2_consts.csv python tools/profile_small.py 8192 0xFF > $O/s2_consts_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 -o $O/s2_fast_full -f python tools/profile_small.py 8192 0xFF > $O/s2_ncu_full.log 2>&1
RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_ls16.so ncu --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 -o $O/s2_ls16_warpstate -f python tools/profile_small.py 8192 0xFF > $O/s2_ncu_ls16.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s2_launches.csv python bench.py --steps 2 --warmup 1 --points 65536 --no-cpu-baseline --no-parity > $O/s2_ncu_launches.log 2>&1
tail -3 $O/s2_pytest.log; cat $O/s2_bench_c3.json | cut -c1-1500
