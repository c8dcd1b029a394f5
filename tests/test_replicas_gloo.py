"""N>1 host logic on CPU: world_size-2 gloo run of the replica partition + max-over-ranks timing
reduction that bench.py uses under torchrun (no data-path collective exists to test)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from upsnet_b200 import replicas as R
    mine = R.image_indices_for_rank(7, rank, world)
    R.barrier()
    t = R.max_over_ranks(10.0 + rank)            # rank 1 is "slower"
    counts = R.gather_counts(len(mine))
    q.put((rank, mine, t, counts))
    dist.destroy_process_group()


def test_two_rank_partition_and_max_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    assert sorted(res[0][1] + res[1][1]) == list(range(7))      # every image exactly once
    assert res[0][2] == res[1][2] == 11.0                        # max over ranks, same on every rank
    assert res[0][3] == res[1][3] == [4, 3]


def test_single_process_degenerates():
    from upsnet_b200 import replicas as R
    assert R.image_indices_for_rank(3, 0, 1) == [0, 1, 2]
    assert R.max_over_ranks(5.5) == 5.5 and R.gather_counts(3) == [3]
