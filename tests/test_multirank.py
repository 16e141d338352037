"""The N > 1 path of bench.py on CPU: two gloo ranks, max-over-ranks timing and whole-job
aggregation (no data-path collective exists: ensemble members are independent)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import bench

    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank r "measured" (10 + 5 r) ms for K = 4 evaluations
    total_ms = bench.max_over_ranks([10.0 + 5.0 * rank, 1.0 + rank], device="cpu")
    value = bench.whole_job_rate(world, 4, total_ms[0])
    seeds = bench.replica_seed(rank)
    dist.barrier()
    out.put((rank, total_ms, value, seeds))
    dist.destroy_process_group()


def test_two_rank_aggregation():
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total_ms, value, seed in got:
        assert total_ms == [15.0, 2.0]                # max over ranks, on every rank
        assert value == pytest.approx(2 * 4 / 15.0e-3)  # all ranks' evaluations / slowest rank's time
        assert seed == rank                            # replica r is the ensemble member with SHUFFLING_SEED = r
