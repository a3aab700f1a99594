"""Frame-level sharding across the GPUs of one box (BASELINE config 4).

Frames are independent end to end (per-sample hash / grid index, per-sample window lists), so the
backbone shards by frame with NO data-path collective: rank r of `world` takes frames
r, r + world, ... .  torch.distributed is only used for the barrier around a timed region and the
max-over-ranks reduction of the measured time (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def frames_for_rank(num_frames, rank, world):
    """round-robin shard: indices of the frames rank `rank` processes"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, num_frames, world))


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (identity without an initialised process group)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


def aggregate_throughput(units_this_rank, seconds_this_rank, device="cpu"):
    """whole-job throughput = units processed by all ranks / max time over ranks"""
    total = sum_over_ranks(units_this_rank, device)
    slowest = max_over_ranks(seconds_this_rank, device)
    return total / slowest
