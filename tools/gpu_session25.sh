#!/bin/bash
# GPU session: the full validation of the head build (tests, constants from ncu, every BASELINE config, ncu captures)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
P=s25
timeout 1500 python -m pytest tests -m gpu -q > $O/${P}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${P}_pytest.log
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 200 python tools/profile_small.py 8192 0xFF > $O/${P}_plain_run.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 --csv --log-file $O/${P}_consts.csv python tools/profile_small.py 8192 0xFF > $O/${P}_consts_run.log 2>&1
cp $O/${P}_consts.csv profiles/r02_consts.csv; cp $O/${P}_consts_run.log profiles/r02_consts_run.log
python tools/ncu_constants.py profiles/r02_consts.csv profiles/r02_consts_run.log profiles/kernel_constants.json > $O/${P}_constants.log 2>&1
cp profiles/kernel_constants.json $O/${P}_kernel_constants.json
timeout 400 python bench.py --steps 5 --warmup 3 > $O/${P}_bench_c3.json 2> $O/${P}_bench_c3.err
timeout 400 python bench.py --config powerlaw --points 1000000 --steps 2 --warmup 1 > $O/${P}_bench_c2.json 2> $O/${P}_bench_c2.err
timeout 200 python bench.py --config juettner_sweep --steps 5 --warmup 3 > $O/${P}_bench_c5.json 2> $O/${P}_bench_c5.err
timeout 600 python bench.py --config pitchy_kappa --points 32768 --steps 1 --warmup 1 > $O/${P}_bench_c4.json 2> $O/${P}_bench_c4.err
timeout 200 python bench.py --single-process --gpus 1 --points 262144 --steps 2 --warmup 1 --no-parity > $O/${P}_bench_sp1.json 2> $O/${P}_bench_sp1.err
timeout 120 python tools/variant_bench.py 131072 pitchy_pl 2 > $O/${P}_kernels.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 --csv --log-file $O/${P}_dram_65536.csv python tools/profile_small.py 65536 0xFF > $O/${P}_dram_65536_run.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 -o $O/${P}_fast_full -f python tools/profile_small.py 8192 0xFF > $O/${P}_ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches.csv python bench.py --steps 2 --warmup 1 --points 65536 --no-cpu-baseline --no-parity > $O/${P}_ncu_launches.log 2>&1
timeout 300 python tools/kappa_check.py > $O/${P}_kappa.log 2>&1
tail -4 $O/${P}_pytest.log | cut -c1-300; cat $O/${P}_kernels.log $O/${P}_constants.log; cat $O/${P}_kappa.log | head -3
for c in c3 c2 c5 c4 sp1; do python -c "
import json; d=json.load(open('$O/${P}_bench_$c.json')); print('$c', round(d['value']), round(d['e2e']['value']), d['roofline'].get('frac'), d.get('parity',{}).get('meets_north_star'), d.get('parity',{}).get('meets_north_star_but_for_reference_failures'), d.get('cpu_baseline',{}).get('value'))"; done
