"""CPU, world_size 2 over gloo: the frame sharding of the multi-GPU path (no data-path collective;
only barrier + max-over-ranks timing) partitions the frames and aggregates throughput correctly."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mssvt_b200.sharding import aggregate_throughput, frames_for_rank, max_over_ranks


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = frames_for_rank(16, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    dist.barrier()
    seconds = 1.0 + rank  # rank 1 is the slow one
    agg = aggregate_throughput(len(mine) * 150000, seconds)
    slow = max_over_ranks(seconds)
    if rank == 0:
        out.put((gathered, agg, slow))
    dist.destroy_process_group()


def test_frames_are_partitioned_and_time_is_max_over_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, agg, slow = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = sorted(i for part in gathered for i in part)
    assert flat == list(range(16))                      # every frame exactly once
    assert gathered[0] == list(range(0, 16, 2)) and gathered[1] == list(range(1, 16, 2))
    assert slow == 2.0                                   # max over ranks, not the mean
    assert abs(agg - 16 * 150000 / 2.0) < 1e-6           # whole-job units / slowest rank


def test_single_process_is_identity():
    assert frames_for_rank(5, 0, 1) == [0, 1, 2, 3, 4]
    assert max_over_ranks(3.5) == 3.5
