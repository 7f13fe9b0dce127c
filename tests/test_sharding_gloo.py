"""The N > 1 path of bench.py on CPU: two gloo ranks, one shard each, no data-path
collective; only the max-over-ranks of the timing is exchanged."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from rimphony_b200.sampler import synthetic_batch

    kind, s, theta, params = synthetic_batch("pitchy_pl", 1000, seed=bench.SEED, shard=rank)
    # each rank "times" its own shard; the job time is the slowest rank
    fake_ms = [100.0 + 50.0 * rank, 300.0 - 25.0 * rank]
    got = bench.max_over_ranks(fake_ms, world, torch.device("cpu"))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), s=s, theta=theta, p=params[0], got=np.array(got))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduction(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    # both ranks agree on the max over ranks
    assert r0["got"].tolist() == [150.0, 300.0] and r1["got"].tolist() == [150.0, 300.0]
    # shards are different draws of the same distribution; no point is shared
    assert len(r0["s"]) == len(r1["s"]) == 1000
    assert not np.intersect1d(r0["s"], r1["s"]).size
    assert abs(np.log(r0["s"]).mean() - np.log(r1["s"]).mean()) < 0.6
    import bench
    assert bench.whole_job_value(1000, 2, 3, 150.0) == 1000 * 2 * 3 / 0.15
