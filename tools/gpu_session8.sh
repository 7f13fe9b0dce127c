#!/bin/bash
# GPU session: the full validation of the head build (tests, every BASELINE config, ncu captures)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/s8_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s8_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/s8_bench_c3.json 2> $O/s8_bench_c3.err
timeout 400 python bench.py --config powerlaw --points 1000000 --steps 2 --warmup 1 > $O/s8_bench_c2.json 2> $O/s8_bench_c2.err
timeout 200 python bench.py --config juettner_sweep --steps 5 --warmup 3 > $O/s8_bench_c5.json 2> $O/s8_bench_c5.err
timeout 600 python bench.py --config pitchy_kappa --points 32768 --steps 1 --warmup 1 > $O/s8_bench_c4.json 2> $O/s8_bench_c4.err
timeout 200 python bench.py --single-process --gpus 1 --points 262144 --steps 2 --warmup 1 --no-parity > $O/s8_bench_sp1.json 2> $O/s8_bench_sp1.err
timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2 > $O/s8_kernels.log 2>&1
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 200 ncu --metrics $M --clock-control none -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 --csv --log-file $O/s8_consts.csv python tools/profile_small.py 8192 0xFF > $O/s8_consts_run.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 --csv --log-file $O/s8_dram_65536.csv python tools/profile_small.py 65536 0xFF > $O/s8_dram_65536_run.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 -o $O/s8_fast_full -f python tools/profile_small.py 8192 0xFF > $O/s8_ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s8_launches.csv python bench.py --steps 2 --warmup 1 --points 65536 --no-cpu-baseline --no-parity > $O/s8_ncu_launches.log 2>&1
{
timeout 200 python tools/kappa_check.py
for v in sf5 sf10; do RIMPHONY_B200_LIB=$PWD/rimphony_b200/variants/librimphony_b200_$v.so timeout 200 python tools/kappa_check.py; done
} > $O/s8_kappa.log 2>&1
cat $O/s8_kappa.log
tail -12 $O/s8_pytest.log | cut -c1-300; cat $O/s8_kernels.log
