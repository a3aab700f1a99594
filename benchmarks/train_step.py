"""BASELINE config 5: MsSVT backbone train step (forward + backward + AdamW) on synthetic frames,
one frame per GPU per step, DDP gradient all-reduce over NCCL when launched under torchrun.

    python benchmarks/train_step.py [--steps 10] [--warmup 3] [--voxels 150000] [--amp]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29533 benchmarks/train_step.py --amp

Training runs the ragged path of the blocks (default): index maps from the fused geometry kernels, the window
attention and the three-NN blend forward + backward in the hand-written kernels of csrc/train.cu on compact window
lists, the projections / FFN as torch GEMMs (bf16 autocast with --amp).  --path padded = torch autograd over the
reference's padded tensors (the cross-check path), for comparison.
Loss = mean of the squared `.dense()` output (synthetic).  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mssvt_b200.config import s0_model_cfg  # noqa: E402
from mssvt_b200 import mssvt_backbone  # noqa: E402
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer  # noqa: E402
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--voxels", type=int, default=150000)
    ap.add_argument("--amp", action="store_true", help="bf16 autocast for the dense math")
    ap.add_argument("--tf32", action="store_true", help="allow TF32 tensor-core matmuls in torch (default: fp32 SIMT)")
    ap.add_argument("--passes", type=int, default=3, help="timed passes of --steps steps; the median is reported")
    ap.add_argument("--batch", type=int, default=1, help="frames per GPU per step (batch_size of the forward)")
    ap.add_argument("--path", choices=("ragged", "padded"), default="ragged", help="training path of the blocks")
    ap.add_argument("--profile", default=None, help="write a torch.profiler kernel table of two steps to this file")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    mssvt_backbone.TRAIN_PATH = args.path
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(s0_model_cfg(), 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).to(dev)
    model.train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    frames = []
    for i in range(4):
        fs, cs = [], []
        for b in range(args.batch):           # samples of a step: concatenated, batch index in column 0
            f, c = synth_frame(1000 * rank + args.batch * i + b, args.voxels)
            c = c.copy()
            c[:, 0] = b
            fs.append(torch.from_numpy(f))
            cs.append(torch.from_numpy(c))
        frames.append((torch.cat(fs).to(dev), torch.cat(cs).to(dev)))

    def step(i):
        f, c = frames[i % len(frames)]
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.amp):
            sp = net({"voxel_features": f, "voxel_coords": c, "batch_size": args.batch})["encoded_spconv_tensor"]
            loss = (sp.dense().float() ** 2).mean()
        loss.backward()
        opt.step()
        return loss

    for i in range(args.warmup):
        step(i)
    # `passes` timed passes of `steps` steps each; the median is reported (at one frame per step the step is bound by the
    # host, and a busy host shows up as a slow pass)
    passes, wall = [], 0.0
    for p in range(args.passes):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(args.steps):
            loss = step(args.warmup + i)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        passes.append(ms)
    ms = sorted(passes)[len(passes) // 2]
    # the same steps without the gradient all-reduce (DDP no_sync): the difference is what the NCCL all-reduce
    # costs after overlap with the backward pass
    ms_nosync = None
    if world > 1:
        with net.no_sync():
            for i in range(2):
                step(i)
            torch.cuda.synchronize()
            dist.barrier()
            n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0.record()
            for i in range(args.steps):
                step(args.warmup + i)
            n1.record()
            torch.cuda.synchronize()
        t = torch.tensor([n0.elapsed_time(n1) / args.steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_nosync = float(t.item())
    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for i in range(2):
                step(i)
            torch.cuda.synchronize()
        with open(args.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
    if rank == 0:
        grad_bytes = sum(p.numel() for p in model.parameters()) * 4
        print(json.dumps({"metric": "mssvt_backbone_train_step_ms", "value": ms, "unit": "ms/step", "n_gpus": world,
                          "ms_per_step_without_allreduce": ms_nosync,
                          "allreduce_share": None if ms_nosync is None else max(0.0, (ms - ms_nosync) / ms),
                          "gradient_bytes": grad_bytes,
                          "steps": args.steps, "warmup": args.warmup, "higher_is_better": False,
                          "voxels_per_s": args.voxels * args.batch * world / (ms * 1e-3), "frames_per_gpu_per_step": args.batch, "wall_ms_per_step": wall / args.steps * 1e3,
                          "dtype": "bf16 autocast" if args.amp else "tf32 matmuls" if args.tf32 else "fp32", "loss": float(loss.detach()), "train_path": args.path, "passes_ms_per_step": [round(p, 3) for p in passes],
                          "config": {"workload": "S0 backbone fwd+bwd+AdamW, %d synthetic %d-voxel frame(s) per GPU per "
                                                 "step, loss = mean(dense()^2), DDP all-reduce when world > 1" % (args.batch, args.voxels)}}),
              file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
