"""Window-set sharding of ONE large frame across the GPUs of a box (second row).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
        benchmarks/shard_frame.py [--voxels 1000000] [--steps 10]

Every rank holds the full frame, takes an x-slab (+ 1-voxel halo), runs the unmodified kernels on it and
exchanges only the halo rows with its neighbours after every attention block (mssvt_b200/sharding.py).  The
script (1) checks that the rows a rank owns equal the rows of the single-GPU forward BIT FOR BIT, (2) times
single-GPU against sharded (max over ranks, CUDA events).  Prints one JSON line."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mssvt_b200.config import s0_model_cfg  # noqa: E402
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer  # noqa: E402
from mssvt_b200.sharding import sharded_backbone_forward  # noqa: E402
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voxels", type=int, default=1000000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--precision", default="bf16x3")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = s0_model_cfg()
    cfg["PRECISION"] = args.precision
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).to(dev).eval()
    for b in model.backbone:
        b.max_num_wins = max(b.max_num_wins, args.voxels)       # (the reference's 90 000-window cap is per sample)
    f, c = synth_frame(77, args.voxels)
    f, c = torch.from_numpy(f).to(dev), torch.from_numpy(c).to(dev)

    def single():
        with torch.no_grad():
            return model({"voxel_features": f, "voxel_coords": c, "batch_size": 1})["encoded_spconv_tensor"]

    sorted_x = bool((c[1:, 3] >= c[:-1, 3]).all())     # (checked once, outside the timed region)

    def sharded():
        return sharded_backbone_forward(model, f, c, 1, rank, world, sorted_by_x=sorted_x)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    ref = single()
    ref_f, ref_i = ref.features, ref.indices
    got_f, got_i, plan = sharded()
    # match the owned pillars to the single-GPU rows by pillar key
    key = lambda i: (i[:, 0].long() * 100000 + i[:, 3].long()) * 100000 + i[:, 2].long()
    order_ref = torch.argsort(key(ref_i))
    kr = key(ref_i)[order_ref]
    pos = torch.searchsorted(kr, key(got_i))
    same_rows = bool((kr[pos] == key(got_i)).all())
    equal = bool(torch.equal(ref_f[order_ref][pos], got_f))
    maxdiff = float((ref_f[order_ref][pos] - got_f).abs().max()) if got_f.numel() else 0.0
    counts = torch.tensor([got_f.shape[0]], device=dev)
    if world > 1:
        dist.all_reduce(counts)
    covered = int(counts.item()) == ref_f.shape[0]
    flags = torch.tensor([int(same_rows and equal)], device=dev)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    t_single, t_sharded = timed(single), timed(sharded)
    marks = []
    torch.cuda.synchronize()
    sharded_backbone_forward(model, f, c, 1, rank, world, marks=marks, sorted_by_x=sorted_x)
    torch.cuda.synchronize()
    stages = {b[0]: round(a[1].elapsed_time(b[1]), 3) for a, b in zip(marks, marks[1:])}
    if rank == 0:
        print(json.dumps({"metric": "mssvt_one_frame_window_set_sharding_ms", "n_gpus": world, "voxels": args.voxels,
                          "single_gpu_ms": t_single, "sharded_ms": t_sharded, "speedup": t_single / t_sharded,
                          "bit_identical_on_all_ranks": bool(flags.item()), "all_output_rows_covered": covered,
                          "max_abs_diff_rank0": maxdiff, "slab_borders_x": plan.bounds, "halo_voxels": plan.halo,
                          "halo_rows_rank0": int(plan.recv_left.shape[0] + plan.recv_right.shape[0]),
                          "local_rows_rank0": int(plan.local_rows.shape[0]), "scaling": "strong", "rows_sorted_by_x": sorted_x, "stage_ms_rank0": stages,
                          "precision": args.precision}), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
