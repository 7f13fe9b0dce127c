#!/bin/bash
# Multi-GPU session (run with gpurun --gpus 8): the real sharding path and the weak-scaling bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/m_gpus.log 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_entry or current_device or async_calls" > $O/m_pytest.log 2>&1; echo "rc=$?" >> $O/m_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > $O/m_bench_weak8.json 2> $O/m_bench_weak8.err
timeout 300 python bench.py --single-process --gpus 8 --points 10000000 --steps 1 --warmup 1 --no-parity > $O/m_bench_sp8_10M.json 2> $O/m_bench_sp8_10M.err
timeout 300 python bench.py --single-process --gpus 4 --points 10000000 --steps 1 --warmup 1 --no-parity > $O/m_bench_sp4_10M.json 2> $O/m_bench_sp4_10M.err
timeout 200 python bench.py --single-process --gpus 8 --points 2097152 --steps 2 --warmup 1 --no-parity > $O/m_bench_sp8_2M.json 2> $O/m_bench_sp8_2M.err
timeout 200 python bench.py --single-process --gpus 2 --points 2097152 --steps 1 --warmup 1 --no-parity > $O/m_bench_sp2_2M.json 2> $O/m_bench_sp2_2M.err
tail -3 $O/m_pytest.log; for f in weak8 sp8_10M sp4_10M sp8_2M sp2_2M; do python -c "
import json,sys
try:
    d=json.load(open('$O/m_bench_$f.json')); print('$f', d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['scaling'])
except Exception as e: print('$f failed', e)
"; done
