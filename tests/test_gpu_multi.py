"""Real multi-GPU test (needs >= 2 GPUs: gpurun --gpus 2): one system, pair work sharded over two
processes with NCCL, against the single-GPU result — bit-identical by construction."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rank(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch
    import torch.distributed as dist
    from common import make_case, to_engine
    from multimm_b200.engine import Engine

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    box = [Engine.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    case = make_case(20000, n_chrom=5, seed=3)
    eng = to_engine(case, device=rank)
    eng.dist_init(rank, world, box[0])
    e, f = eng.energy_forces()
    rep = eng.minimize(tol=10.0, max_iter=20)
    x = eng.get_positions()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, e, f, rep, x))


def test_two_gpus_one_system(built_lib):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from common import make_case, to_engine

    case = make_case(20000, n_chrom=5, seed=3)
    eng = to_engine(case, device=0)
    e0, f0 = eng.energy_forces()
    rep0 = eng.minimize(tol=10.0, max_iter=20)
    x0 = eng.get_positions()
    eng.close()

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank, args=(r, 2, 29741, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e, f, rep, x in got:
        assert np.array_equal(e, e0) and np.array_equal(f, f0)
        assert rep["e_final"] == rep0["e_final"] and rep["evaluations"] == rep0["evaluations"]
        assert np.array_equal(x, x0)
