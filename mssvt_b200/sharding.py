"""Frame-level sharding across the GPUs of one box (BASELINE config 4).

Frames are independent end to end (per-sample hash / grid index, per-sample window lists), so the
backbone shards by frame with NO data-path collective: rank r of `world` takes frames
r, r + world, ... .  torch.distributed is only used for the barrier around a timed region and the
max-over-ranks reduction of the measured time (NCCL on GPUs, gloo in the CPU tests)."""
import math

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


# ---------------------------------------------------------------------------------------------------------
# Window-set sharding of ONE large frame (second row; north star: "by window sets within a
# large frame").  Domain decomposition instead of an all-gather: the frame is cut into x-slabs whose borders
# are window-aligned, every rank runs the unmodified kernels on its slab plus a halo as wide as the larger
# window sticks out of the smaller one (1 voxel for 3^3 / 5^3), and after every attention block only the halo
# rows are exchanged with the two neighbours (point-to-point, ~1 % of the rows).  The windows of an owned voxel
# and the key candidates of those windows lie inside slab + halo, so owned rows are computed exactly as on one
# GPU; rows in the halo are recomputed by their owner and overwritten by the exchange.
# One more row travels: the reference lets FPS pick a padded slot, and that key aliases voxel 0 of the sample
# ((-1 + 0.1).int() == 0, quirk Q1) WITHOUT being masked, so every window with padding may read the features of
# the sample's first voxel.  Each rank therefore keeps the first voxel of every sample in its local frame (it
# stays local row 0 of the sample) and refreshes it from its owner after every block.


class SlabPlan:
    """x-slab decomposition of one frame.  Every rank holds the full coordinate list (inputs are replicated),
    so all index lists are derived locally and neighbours agree on sizes and row order without a handshake."""

    def __init__(self, coords, win_x, halo, rank, world, grid_x=None, sorted_single_sample=False):
        """coords (N, 4) int [b, z, y, x]; win_x: x extent of the window grid the slab borders align to;
        halo: voxels a slab needs beyond its borders; grid_x: x extent of the voxel grid (saves a host sync);
        sorted_single_sample: the caller guarantees ONE sample whose rows ascend in x (the order DynamicVFE /
        torch.unique produce): every row set below is then a contiguous range found by binary search"""
        x = coords[:, 3]
        n = x.shape[0]
        if sorted_single_sample and n and world > 1:
            self._init_sorted(x, n, win_x, halo, rank, world, grid_x)
            return
        if grid_x is None:
            grid_x = int(x.max().item()) + 1 if n else 1
        hist = torch.bincount(x, minlength=grid_x)
        cum = torch.cumsum(hist, 0)
        # equal voxel counts, borders on multiples of win_x; one readback for all cuts
        targets = torch.tensor([n * r // world for r in range(1, world)], device=cum.device, dtype=cum.dtype)
        cuts = (torch.searchsorted(cum, targets) + 1).tolist() if world > 1 else []
        bounds = [0]
        for cut in cuts:
            bounds.append(max(bounds[-1], (cut + win_x // 2) // win_x * win_x))
        bounds.append(hist.shape[0] + win_x)
        self._check_bounds(bounds, halo, world)
        self.bounds, self.rank, self.world, self.halo = bounds, rank, world, halo
        lo, hi = bounds[rank], bounds[rank + 1]
        self.lo, self.hi = lo, hi
        # slab + halo is one x range; plus the first voxel of every sample (samples are contiguous and ordered)
        sel = (x >= (lo - halo if rank > 0 else lo)) & (x < (hi + halo if rank < world - 1 else hi))
        nb = int(coords[-1, 0].item()) + 1 if n else 0
        starts = torch.searchsorted(coords[:, 0].contiguous(), torch.arange(nb, device=x.device, dtype=coords.dtype))
        sel[starts] = True
        first = torch.zeros_like(sel)
        first[starts] = True
        self.local_rows = torch.nonzero(sel).squeeze(1)                      # ascending: global row order kept
        xl = x[self.local_rows]
        fl = first[self.local_rows]
        none = torch.zeros_like(fl)
        # rows of the LOCAL tensor: what the neighbours need from me, where their rows land here, what I own,
        # the samples' first voxels -- six masks, one nonzero, one readback of the six counts
        masks = torch.stack([
            (xl >= lo) & (xl < lo + halo) if rank > 0 else none,
            (xl >= hi - halo) & (xl < hi) if rank < world - 1 else none,
            (xl >= lo - halo) & (xl < lo) if rank > 0 else none,
            (xl >= hi) & (xl < hi + halo) if rank < world - 1 else none,
            (xl >= lo) & (xl < hi),
            fl])
        nz = torch.nonzero(masks)[:, 1]
        counts = masks.sum(1).tolist()
        parts = torch.split(nz, counts)
        self.send_left, self.send_right, self.recv_left, self.recv_right, self.owned_local, self.alias_local = parts
        self.alias_mine = ((xl >= lo) & (xl < hi))[self.alias_local]          # which first voxels this rank owns

    @staticmethod
    def _check_bounds(bounds, halo, world):
        """every rank derives the same bounds, so every rank raises together: a slab narrower than the halo
        (or empty) would make the send / receive row sets of neighbours disagree"""
        if world > 1 and any(b - a < max(halo, 1) for a, b in zip(bounds, bounds[1:])):
            raise ValueError("one-frame sharding: a slab is narrower than the halo (%d voxels) -- slab borders %s; "
                             "use fewer ranks for this frame" % (halo, bounds))

    def _init_sorted(self, x, n, win_x, halo, rank, world, grid_x):
        dev = x.device
        pick = torch.tensor([n * r // world for r in range(1, world)], device=dev)
        cuts = (x[pick] + 1).tolist()                                   # read-back 1: the quantile columns
        bounds = [0]
        for cut in cuts:
            bounds.append(max(bounds[-1], (cut + win_x // 2) // win_x * win_x))
        bounds.append((grid_x if grid_x is not None else int(x[-1].item()) + 1) + win_x)
        self._check_bounds(bounds, halo, world)
        self.bounds, self.rank, self.world, self.halo = bounds, rank, world, halo
        lo, hi = bounds[rank], bounds[rank + 1]
        self.lo, self.hi = lo, hi
        has_l, has_r = rank > 0, rank < world - 1
        marks = torch.tensor([lo - halo if has_l else lo, lo, lo + halo, hi - halo, hi, hi + halo if has_r else hi],
                             device=dev, dtype=x.dtype)
        p = torch.searchsorted(x.contiguous(), marks).tolist()          # read-back 2: six row positions
        start, end = p[0], p[5]
        extra = 1 if start > 0 else 0                                   # the sample's first voxel rides in front
        rows = torch.arange(start, end, device=dev)
        self.local_rows = torch.cat([rows.new_zeros(1), rows]) if extra else rows
        rng = lambda a, b: torch.arange(a - start + extra, b - start + extra, device=dev)
        none = rows[:0]
        self.send_left = rng(p[1], p[2]) if has_l else none
        self.send_right = rng(p[3], p[4]) if has_r else none
        self.recv_left = rng(p[0], p[1]) if has_l else none
        self.recv_right = rng(p[4], p[5]) if has_r else none
        self.owned_local = rng(p[1], p[4])
        self.alias_local = rows.new_zeros(1)
        self.alias_mine = torch.tensor([p[1] == 0 and p[4] > 0], device=dev)
        # the same row sets as (first, last) ranges of the local tensor: every set is contiguous here, so the local
        # frame is a slice of the inputs and the halo rows travel from / into views (no gather, no scatter)
        off = lambda a, b: (a - start + extra, b - start + extra)
        self.global_range = (start, end, extra)
        self.ranges = {"send_left": off(p[1], p[2]) if has_l else None, "send_right": off(p[3], p[4]) if has_r else None,
                       "recv_left": off(p[0], p[1]) if has_l else None, "recv_right": off(p[4], p[5]) if has_r else None}

    def exchange_many(self, tensors, group=None):
        """exchange() for several (n_local, C) tensors at once: ONE batch of point-to-point transfers and ONE
        all-reduce for the lot (the y rows and the next LayerNorm rows of a block travel together).  With the
        contiguous row sets of a frame sorted in x the halo rows are sent from and received into slices of the
        tensors themselves."""
        tensors = [t for t in tensors if t is not None]
        if self.world == 1 or not tensors:
            return tensors
        ranges = getattr(self, "ranges", None)
        if ranges is None:
            for t in tensors:
                self.exchange(t, group)
            return tensors
        ops = []
        for peer, send, recv in ((self.rank - 1, ranges["send_left"], ranges["recv_left"]),
                                 (self.rank + 1, ranges["send_right"], ranges["recv_right"])):
            if peer < 0 or peer >= self.world or send is None:
                continue
            for t in tensors:
                if send[1] > send[0]:
                    ops.append(dist.P2POp(dist.isend, t[send[0]:send[1]], peer, group))
                if recv[1] > recv[0]:
                    ops.append(dist.P2POp(dist.irecv, t[recv[0]:recv[1]], peer, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        # the samples' first voxels (local row 0 here): exactly one rank contributes a non-zero row
        alias = torch.stack([t[0] for t in tensors]) * self.alias_mine.to(tensors[0].dtype)
        dist.all_reduce(alias, group=group)
        for i, t in enumerate(tensors):
            t[0].copy_(alias[i])
        return tensors

    def exchange(self, features, group=None):
        """overwrite the halo rows of `features` (n_local, C) with the owners' values"""
        if self.world == 1:
            return features
        ops, bufs = [], []
        for peer, send_idx, recv_idx in ((self.rank - 1, self.send_left, self.recv_left),
                                         (self.rank + 1, self.send_right, self.recv_right)):
            if peer < 0 or peer >= self.world:
                continue
            out = features.index_select(0, send_idx).contiguous()
            buf = features.new_empty((recv_idx.shape[0], features.shape[1]))
            ops += [dist.P2POp(dist.isend, out, peer, group), dist.P2POp(dist.irecv, buf, peer, group)]
            bufs.append((recv_idx, buf, out))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for recv_idx, buf, _ in bufs:
            features.index_copy_(0, recv_idx, buf)
        # the samples' first voxels: exactly one rank contributes a non-zero row, the sum is that row
        alias = features.index_select(0, self.alias_local) * self.alias_mine.unsqueeze(1).to(features.dtype)
        dist.all_reduce(alias, group=group)
        features.index_copy_(0, self.alias_local, alias)
        return features


def sharded_backbone_forward(model, voxel_features, voxel_coords, batch_size, rank, world, group=None, marks=None,
                             sorted_by_x=False):
    """One frame over `world` ranks.  Inputs are the FULL frame on every rank; returns (features, indices) of
    the output rows this rank owns (the pillars of its slab), identical to the corresponding rows of the
    single-GPU forward.  Inference only.  sorted_by_x: the caller guarantees rows ascending in x within the
    (single) sample, which makes the slab plan two binary searches instead of passes over all voxels."""
    from .mssvt_backbone import MixedScaleSparseTransformerCompressBlock as Compress

    def mark(name):                    # optional stage timeline: (name, CUDA event) pairs appended to `marks`
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    coords = voxel_coords if voxel_coords.dtype == torch.int32 else voxel_coords.int()
    blocks = list(model.backbone)
    # slab borders must lie on the window grid of EVERY block: the least common multiple of the x extents
    win_x = 1
    for b in blocks:
        win_x = math.lcm(win_x, int(b.win1_size[0]))
    halo = max([(int(b.win2_size[0]) - int(b.win1_size[0]) + 1) // 2 for b in blocks if b.win2_size is not None] + [0])
    plan = SlabPlan(coords, win_x, halo, rank, world, grid_x=int(model.grid_size[0]),
                    sorted_single_sample=bool(sorted_by_x) and batch_size == 1)
    span = getattr(plan, "global_range", None)
    if span is not None and not span[2]:          # a slice of the inputs (nothing rides in front)
        local_f, local_c = voxel_features[span[0]:span[1]], coords[span[0]:span[1]]
    elif span is not None:                        # the sample's first voxel + a slice
        local_f = torch.cat([voxel_features[:1], voxel_features[span[0]:span[1]]])
        local_c = torch.cat([coords[:1], coords[span[0]:span[1]]])
    else:
        local_f = voxel_features.index_select(0, plan.local_rows)
        local_c = coords.index_select(0, plan.local_rows)
    sp = model._sparse_tensor(local_f.contiguous(), local_c.contiguous(), batch_size)
    mark("plan + local frame")
    with torch.no_grad():
        for i, block in enumerate(blocks):
            sp = block(sp, block_idx=i)
            mark("block %d" % i)
            if isinstance(block, Compress):
                break
            pre = getattr(sp, "_xn_ready", None)   # the next block's LayerNorm rows ride along (written by the FFN epilogue)
            plan.exchange_many([sp.features, pre[1] if pre is not None else None], group)
            mark("exchange %d" % i)
        feats, idx = sp.features, sp.indices
        if isinstance(blocks[-1], Compress) and span is not None:
            # rows ascend in x, windows are listed by first occurrence: the owned output rows are one range
            wx = int(blocks[-1].win1_size[0])
            px = (idx[:, 3] * wx).contiguous()
            a, b = torch.searchsorted(px, torch.tensor([plan.lo, plan.hi], device=px.device, dtype=px.dtype)).tolist()
            out = feats[a:b], idx[a:b], plan
            mark("owned rows")
            return out
        if isinstance(blocks[-1], Compress):
            wx = int(blocks[-1].win1_size[0])   # output rows are windows of the compress grid
            keep = (idx[:, 3].long() * wx >= plan.lo) & (idx[:, 3].long() * wx < plan.hi)
        else:
            keep = torch.zeros(feats.shape[0], dtype=torch.bool, device=feats.device)
            keep[plan.owned_local] = True
        out = feats[keep], idx[keep], plan
        mark("owned rows")
        return out
