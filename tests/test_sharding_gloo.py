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


# ---- window-set (x-slab) sharding of one frame: plan + halo exchange over gloo -----------------------------

def _slab_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mssvt_b200.sharding import SlabPlan
    from mssvt_b200.synth import synth_frame
    _, coords = synth_frame(3, 6000, crop=0.25)
    coords = torch.from_numpy(coords)
    plan = SlabPlan(coords, 3, 1, rank, world)
    # "features" = global row id in every channel, owned rows marked with the owner's rank in channel 1;
    # halo rows start as garbage and must come back holding the owner's values
    feats = torch.full((plan.local_rows.shape[0], 4), -1.0)
    feats[plan.owned_local, 0] = plan.local_rows[plan.owned_local].float()
    feats[plan.owned_local, 1] = float(rank)
    plan.exchange(feats)
    x = coords[plan.local_rows, 3]
    ok_rows = bool((feats[:, 0] == plan.local_rows.float()).all())          # (incl. the sample's first voxel)
    owner = torch.bucketize(x, torch.tensor(plan.bounds[1:-1]), right=True).float()
    ok_owner = bool((feats[:, 1] == owner).all()) and int(plan.local_rows[plan.alias_local[0]]) == 0
    gathered = [None] * world
    dist.all_gather_object(gathered, (plan.bounds, int(plan.owned_local.shape[0]), ok_rows, ok_owner,
                                      int(plan.recv_left.shape[0] + plan.recv_right.shape[0])))
    if rank == 0:
        out.put((gathered, coords.shape[0]))
    dist.destroy_process_group()


def test_slab_plan_and_halo_exchange_world3():
    world = 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, n = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(g[0] == gathered[0][0] for g in gathered)            # every rank derives the same borders
    assert all(b % 3 == 0 for b in gathered[0][0][1:-1])            # window-aligned
    assert sum(g[1] for g in gathered) == n                         # every voxel owned exactly once
    assert all(g[2] and g[3] for g in gathered)                     # halo rows hold their owners' values
    assert all(g[4] > 0 for g in gathered)


def test_slab_plan_sorted_fast_path_equals_generic_path():
    """frames whose rows ascend in x (the order DynamicVFE / torch.unique produce): the plan from two binary
    searches is the plan from the passes over all voxels"""
    from mssvt_b200.sharding import SlabPlan
    from mssvt_b200.synth import synth_frame
    for seed, n, crop in ((3, 20000, 0.4), (5, 3000, 0.2)):
        _, c = synth_frame(seed, n, crop=crop)
        c = torch.from_numpy(c)
        assert bool((c[1:, 3] >= c[:-1, 3]).all())
        for world in (2, 3, 5):
            for r in range(world):
                a = SlabPlan(c, 3, 1, r, world, grid_x=468)
                b = SlabPlan(c, 3, 1, r, world, grid_x=468, sorted_single_sample=True)
                assert a.bounds == b.bounds
                for name in ("local_rows", "send_left", "send_right", "recv_left", "recv_right", "owned_local",
                             "alias_local"):
                    assert torch.equal(getattr(a, name), getattr(b, name)), (seed, world, r, name)
                assert a.alias_mine.tolist() == b.alias_mine.tolist()


def _slab_many_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mssvt_b200.sharding import SlabPlan
    from mssvt_b200.synth import synth_frame
    _, coords = synth_frame(3, 6000, crop=0.25)
    coords = torch.from_numpy(coords)
    plan = SlabPlan(coords, 3, 1, rank, world, grid_x=468, sorted_single_sample=True)
    # the row ranges of the sorted plan are its index lists
    ok_ranges = all(
        (plan.ranges[k] is None and getattr(plan, k).numel() == 0) or
        (plan.ranges[k] is not None and torch.equal(torch.arange(*plan.ranges[k]), getattr(plan, k)))
        for k in ("send_left", "send_right", "recv_left", "recv_right"))
    start, end, extra = plan.global_range
    ok_ranges = ok_ranges and torch.equal(plan.local_rows[extra:], torch.arange(start, end)) and \
        (extra == 0 or int(plan.local_rows[0]) == 0)
    # two tensors travel in one batch: "y" = global row id, "xn" = 1000000 + global row id, owned rows only
    y = torch.full((plan.local_rows.shape[0], 4), -1.0)
    xn = torch.full((plan.local_rows.shape[0], 4), -1.0)
    y[plan.owned_local] = plan.local_rows[plan.owned_local].float().unsqueeze(1)
    xn[plan.owned_local] = 1000000.0 + plan.local_rows[plan.owned_local].float().unsqueeze(1)
    plan.exchange_many([y, None, xn])
    ok = bool((y[:, 0] == plan.local_rows.float()).all()) and bool((xn[:, 0] == 1000000.0 + plan.local_rows.float()).all())
    gathered = [None] * world
    dist.all_gather_object(gathered, (ok_ranges, ok))
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


def test_sorted_plan_exchanges_several_tensors_from_views_world3():
    """the halo rows of y and of the next LayerNorm rows in ONE batch of point-to-point transfers, sent from and
    received into slices of the tensors (frames sorted in x), plus the samples' first voxels in one all-reduce"""
    world = 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_many_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(g[0] for g in gathered) and all(g[1] for g in gathered), gathered
