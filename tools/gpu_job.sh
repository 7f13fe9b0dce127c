set -x
python tools/profile_small.py 4096 0xFF > gpurun_out/plain_fast4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 -o gpurun_out/prof_fast4 -f python tools/profile_small.py 4096 0xFF > gpurun_out/ncu_fast4.log 2>&1
python bench.py --points 32768 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --points 32768 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1
cat gpurun_out/plain_fast4.log; tail -3 gpurun_out/ncu_fast4.log
