#!/usr/bin/env python
"""bench.py -- MsSVT backbone forward throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm (sm_100a kernels)
    python bench.py --impl reference --gpus N ...            reference arm: the CPU restatement of
                                                             the reference path on the host cores

A step = one backbone forward (3 mixed-scale blocks + z-compress block, config S0) over one
synthetic Waymo-scale frame per GPU (BASELINE config 2; with N > 1 GPUs every rank runs its own
frames, config 4: sharded by frame, no data-path collective, weak scaling).
  value  = voxels/s with the frames already resident in HBM, CUDA-event time over exactly K steps,
           max over ranks;
  e2e    = the same metric through the module API from pinned HOST buffers: H2D of the step's
           features + coordinates, forward, D2H of the output features + indices, every step;
  roofline      = the dominant kernel of the step against the measured HBM peak;
  cpu_baseline  = the CPU oracle on the host cores, bounded sample (rank 0, N = 1 only);
  train_step    = side measurement after everything else (N = 1 only, never part of `value`): one training step
                  (forward + backward + AdamW) through the hand-written training kernels.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mssvt_b200.config import s0_model_cfg  # noqa: E402
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame, synth_points  # noqa: E402

METRIC = "mssvt_backbone_fwd_voxels_per_s"
UNIT = "voxels/s"
N_VOXELS = 150000
POOL = 8  # distinct frames rotated through the timed loop: 8 x 40.8 MB = 326 MB > 126 MB L2


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled from a thread
    every millisecond (the timed region is ~20 ms: `nvidia-smi -lms` does not even start in that time)."""

    def __init__(self, device_index):
        self.idx, self.samples, self.reasons, self.max_mhz = device_index, [], set(), None
        self._stop, self._thread, self._err, self._ready = False, None, None, None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index(nv))
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            self._ready.set()
            while True:          # (at least one sample, however short the timed region is)
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(h))
                self.reasons.update(n for n, bit in names.items() if mask & bit)
                if self._stop:
                    break
                time.sleep(0.001)
        except Exception as e:  # noqa: BLE001
            self._err = repr(e)
        finally:
            self._ready.set()

    def _physical_index(self, nv):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.idx < len(ids) and ids[self.idx].isdigit():
                return int(ids[self.idx])
        return self.idx

    def start(self):
        import threading
        self._ready = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        self._ready.wait(timeout=5.0)   # NVML is initialised and sampling before the timed region opens (a frame loop of
                                        # 20 steps lasts 15 ms: a fixed 10 ms head start lost the race on a slow nvmlInit)

    def stop(self):
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=2)
        sm = self.samples
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable: %s" % self._err],
                    "samples": 0}
        busy = [v for v in sm if self.max_mhz and v > 0.5 * self.max_mhz] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def build_model(device, precision="fp32", cbs_patterns=(1, 1, 1)):
    from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer
    cfg = s0_model_cfg(cbs_patterns=tuple(cbs_patterns))
    cfg["PRECISION"] = precision
    torch.manual_seed(0)  # random-init weights of the S0 architecture (no checkpoints offline)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    return cfg, model.to(device).eval()


def algorithmic_bytes(kernel, n, w, pillars):
    """Algorithmic HBM bytes per launch (DESIGN.md, section 'Kernels'); fp32, C = 64, S0."""
    C4 = 64 * 4
    table = {
        # read xn rows once + maps, write the covered rows of `merged`
        "mssvt_block_attention": n * C4 + w * (12 * 4 + 64 * 4 + 64 + 27 * 4 + 27 * 3 * 5) + n * 12 + n * C4,
        # read xn rows once + compact maps + coordinates, write the projected row of every real query (~N / 2)
        "mssvt_block_attention_tc": n * C4 + w * (12 * 4 + 64 * 4 + 16) + n * 12 + (n // 2) * C4,
        # the tensor-core FFN with the interpolation + merge on the way in: x, three projected rows per covered
        # voxel out of the L2-resident (N / 2, 64) array (counted once), maps; writes y and the next LayerNorm
        "mssvt_ffn_tc": n * C4 + (n // 2) * C4 + n * (4 + 3 + 12) + 2 * n * C4,
        # read x + merged + covered flag, write y
        "mssvt_ffn": 3 * n * C4 + n,
        "mssvt_layernorm": 2 * n * C4,
        # read xn + xyz + rows map, write one row per pillar
        "mssvt_compress_attention": n * C4 + n * 12 + pillars * 32 * 4 + pillars * C4,
        "mssvt_compress_attention_tc": n * C4 + n * 12 + pillars * 32 * 4 + pillars * C4,
        # window rows in, compact maps out; probes hit the L2-resident 3.2 MB table
        "mssvt_block_geometry": w * 16 + w * (12 * 4 + 27 * 4 + 64 * 4 + 64 + 27 * 3 * 5) + n,
        "mssvt_window_partition": 16 * n + 16 * w + 8 * 400000 + 4 * n,
        "mssvt_build_hash_table": 16 * n + 8 * n + 8 * 400000,
    }
    return table.get(kernel)


def bind_to_gpu_numa_node(local):
    """Pin this rank (and therefore its pinned host buffers: first touch) to the NUMA node its GPU hangs off,
    so that host <-> device copies of the 8 ranks do not cross the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:           # nvml pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception as e:  # noqa: BLE001
        print("[bench] NUMA binding skipped: %r" % (e,), file=sys.stderr)
    return None


WORKLOAD = ("full MsSVT backbone forward (S0: 3 mixed-scale blocks 3^3/5^3 windows, 2+2 heads, K=32 + z-compress block), "
            "one synthetic Waymo-scale frame of 150000 voxels per GPU per step, C=64, hash 400000, batch 1; "
            "random-init weights (seed 0)")

PRECISION_NOTE = {
    "fp32": "fp32 FFMA kernels (features within 1e-4 of max|fp32 reference|)",
    "tf32x3": "tcgen05 kernels with split TF32 operands (3xTF32: A_hi W_hi + A_lo W_hi + A_hi W_lo), fp32 accumulate "
              "(features within 1e-4 of max|fp32 reference|): the module default",
    "tf32": "tcgen05 kernels with TF32 operands / fp32 accumulate, rest fp32 (features within 2e-3 of max|fp32 reference|)",
    "bf16x3": "tcgen05 kernels with split bf16 operands (hi + mid, 16 significant bits: A_hi W_hi + A_mid W_hi + A_hi W_mid on "
              "kind::f16), fp32 accumulate; positional embeddings and the compress attention 3xTF32 (features within 1e-4 of "
              "max|fp32 reference|)",
    "bf16": "tcgen05 kernels with bf16 operands (kind::f16) / fp32 accumulate, rest fp32 (features within 2e-2 of "
            "max|fp32 reference|, rms within 5e-3)",
}
PARITY_TOL = {"fp32": 1e-4, "tf32x3": 1e-4, "bf16x3": 1e-4, "tf32": 2e-3, "bf16": 2e-2}
DTYPE = {"fp32": "f32", "tf32": "tf32", "tf32x3": "tf32x3", "bf16x3": "bf16x3", "bf16": "bf16"}


def kernel_traffic():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of an entry point, summed over its
    kernels: profiles/r02_kernel_traffic.json, written by tools/kernel_traffic.py from the ncu launch list of
    `python tools/profile_forward.py` (N = 150 k; ncu flushes caches between kernels: the cold-cache figure)."""
    path = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


class Arm:
    """Everything one rank needs to time the backbone in a precision mode."""

    def __init__(self, args, rank, world, local):
        self.args, self.rank, self.world = args, rank, world
        self.device = torch.device("cuda", local)
        from mssvt_b200 import _lib
        self.lib = _lib
        _lib.load()
        self.cfg, self.model = build_model(self.device, args.precision)
        # synthetic frames: POOL distinct frames per rank (seeds differ per rank: sharded by frame)
        self.host = []
        for i in range(POOL):
            f, c = synth_frame(1000 * rank + i, N_VOXELS)
            self.host.append((torch.from_numpy(f).pin_memory(), torch.from_numpy(c).pin_memory()))
        self.dev = [(f.to(self.device), c.to(self.device).float()) for f, c in self.host]
        self.dev_idx = [c.to(self.device) for _, c in self.host]   # int32 coordinates: the graphs bind to these

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, run, first, count):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        run(first, count)
        e1.record()
        self.barrier()
        return e0.elapsed_time(e1)

    def eager_step(self, i):
        f, c = self.dev[i % POOL]
        return self.model({"voxel_features": f, "voxel_coords": c, "batch_size": 1})["encoded_spconv_tensor"]

    def measure(self, precision, e2e_passes=3, sample_clocks=False, with_points=False):
        """One full measurement in `precision`: K graph-replayed steps on resident frames (device time), the same
        through the module API from pinned host buffers (e2e), the per-entry-point breakdown and rooflines."""
        args, model, device = self.args, self.model, self.device
        model.set_precision(precision)
        out = {}
        with torch.no_grad():
            # one CUDA graph per resident frame (captured once; every replay runs all kernels of the frame,
            # geometry included); --launch eager times the kernel-by-kernel path instead
            graphs = None
            if args.launch == "graph":
                try:
                    graphs = [model.capture({"voxel_features": self.dev[i][0], "voxel_coords": self.dev_idx[i],
                                             "batch_size": 1}) for i in range(POOL)]
                except Exception as e:  # noqa: BLE001 -- a capture problem must not cost the measurement
                    print("[bench] CUDA graph capture failed (%r): timing the eager path" % (e,), file=sys.stderr)
                    graphs = None
                    torch.cuda.synchronize()

            def steps(first, count):
                for i in range(first, first + count):
                    sp = graphs[i % POOL].replay() if graphs is not None else self.eager_step(i)
                return sp

            sp = steps(0, args.warmup)
            pillars = sp.features.shape[0]
            sampler = ClockSampler(device.index) if sample_clocks else None
            if sampler:
                sampler.start()
            launches0 = self.lib.call("mssvt_launch_count")
            ms = self.timed(steps, args.warmup, args.steps)
            launches = self.lib.call("mssvt_launch_count") - launches0
            if graphs is not None:
                launches = sum(graphs[(args.warmup + i) % POOL].launches for i in range(args.steps))
            if sampler:
                out["clocks"] = sampler.stop()
            for i in range(3):
                self.eager_step(i)
            eager_ms = self.timed(lambda a, n: [self.eager_step(i) for i in range(a, a + n)], args.warmup,
                                  args.steps) / args.steps
            e2e_s, e2e_list, rows = self.e2e(graphs is not None, e2e_passes)
            pts = self.e2e(graphs is not None, e2e_passes, from_points=True) if with_points else None
            # per-entry-point breakdown (instrumented extra pass, not part of the timed region)
            self.lib.PROFILE = []
            for i in range(args.steps):
                self.eager_step(args.warmup + i)
            torch.cuda.synchronize()
            per = {}
            for name, a, b in self.lib.PROFILE:
                t = per.setdefault(name, [0.0, 0])
                t[0] += a.elapsed_time(b)
                t[1] += 1
            self.lib.PROFILE = None
            del graphs
        torch.cuda.empty_cache()
        from mssvt_b200.sharding import max_over_ranks
        ms, e2e_s = max_over_ranks(ms, device), max_over_ranks(e2e_s, device)
        total_voxels = N_VOXELS * args.steps * self.world
        if pts is not None:
            pts_s = max_over_ranks(pts[0], device)
            out["e2e_points"] = {
                "value": total_voxels / pts_s, "unit": UNIT, "h2d_bytes_per_step": self.points[0].numel() * 4,
                "d2h_bytes_per_step": pts[2] * 64 * 4 + pts[2] * 4 * 4, "ms_per_step": pts_s / args.steps * 1e3,
                "passes_ms_per_step": pts[1],
                "pipeline": "pinned raw points (180000 x 6 fp32) -> H2D -> DynamicVFE (64 channels; its voxel count is the "
                            "one host sync) -> backbone -> HeightCompression (dense BEV, stays on the device) -> D2H of "
                            "the sparse output rows; median of %d passes of K steps" % len(pts[1])}
        peaks, peak_kind = measured_peaks()
        traffic = kernel_traffic().get(precision, {})
        w_est = int(0.26 * N_VOXELS)
        total_ms = sum(v[0] for v in per.values())
        kernels = {}
        for name, (t_ms, n) in sorted(per.items(), key=lambda kv: -kv[1][0]):
            k = {"ms_per_step": t_ms / args.steps, "launches_per_step": n / args.steps,
                 "share_of_step": t_ms / total_ms}
            ab = algorithmic_bytes(name, N_VOXELS, w_est, pillars)
            if ab:
                gbs = ab / (t_ms / n * 1e-3) / 1e9
                k.update({"algorithmic_bytes_per_launch": ab, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"],
                          "dram_traffic_per_launch": traffic.get(name)})
            kernels[name] = k
        dom = next(iter(kernels))
        d = kernels[dom]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": d.get("achieved_gbs"), "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": d.get("frac_of_hbm_peak"), "traffic": d.get("dram_traffic_per_launch"),
                    "peak_kind": peak_kind, "avg_launch_us": d["ms_per_step"] / d["launches_per_step"] * 1e3,
                    "share_of_step": d["share_of_step"],
                    "algorithmic_bytes_per_launch": d.get("algorithmic_bytes_per_launch")}
        h2d = self.host[0][0].numel() * 4 + self.host[0][1].numel() * 4
        out.update({
            "precision": precision, "dtype": DTYPE[precision], "value": total_voxels / (ms * 1e-3),
            "ms_per_step": ms / args.steps, "eager_ms_per_step": eager_ms, "gpu_launches": int(launches),
            "e2e": {"value": total_voxels / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": rows * 64 * 4 + rows * 4 * 4, "ms_per_step": e2e_s / args.steps * 1e3,
                    "passes_ms_per_step": e2e_list, "note": "median of %d passes of K steps each" % len(e2e_list)},
            "roofline": roofline, "kernels": kernels, "output_rows": int(pillars),
            "note": PRECISION_NOTE[precision]})
        return out

    def alternating_patterns(self, precision):
        """The same frame loop with blocks that alternate their chessboard pattern (cbs_patterns 1, 0, 2: the golden
        configuration).  The default S0 config uses the odd pattern in every block, which lets the three blocks share
        ALL of the coordinate work; with alternating patterns the chessboard probes, FPS passes and key lists are
        still shared and every further pattern costs one mssvt_block_queries launch."""
        args = self.args
        _, model = build_model(self.device, precision, cbs_patterns=(1, 0, 2))
        with torch.no_grad():
            graphs = [model.capture({"voxel_features": self.dev[i][0], "voxel_coords": self.dev_idx[i], "batch_size": 1})
                      for i in range(POOL)]

            def steps(first, count):
                for i in range(first, first + count):
                    graphs[i % POOL].replay()
            steps(0, args.warmup)
            ms = self.timed(steps, args.warmup, args.steps) / args.steps
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        del graphs
        return {"cbs_patterns": [1, 0, 2], "dtype": precision, "ms_per_step": ms,
                "value": N_VOXELS * self.world / (ms * 1e-3), "unit": UNIT,
                "note": "graph replay, device-timed, same frames / steps as the headline"}

    def points_pipeline(self):
        """DynamicVFE (one PFN layer, 64 channels) in front and HeightCompression (plain dense scatter) behind the
        backbone, and pinned raw point clouds whose voxelisation is exactly the POOL frames' coordinates"""
        if not hasattr(self, "vfe"):
            from mssvt_b200.config import AttrDict
            from mssvt_b200.dynamic_vfe import DynamicVFE
            from mssvt_b200.height_compression import HeightCompression
            torch.manual_seed(1)
            self.vfe = DynamicVFE(AttrDict(NUM_FILTERS=[64]), 5, list(S0_VOXEL), list(S0_GRID), list(S0_RANGE)).to(self.device).eval()
            self.bev = HeightCompression(AttrDict(NUM_BEV_FEATURES=64, COMPRESS_LAYER_NUMS=0)).to(self.device).eval()
            self.points = [torch.from_numpy(synth_points(1000 * self.rank + i, N_VOXELS)).pin_memory() for i in range(POOL)]
        return self.vfe, self.bev, self.points

    def e2e(self, graph, passes, from_points=False):
        """module API from pinned HOST buffers; every step copies its inputs host -> device and its result (features
        + indices of the output tensor) device -> host.  Three streams: the H2D of step i+1 and the D2H of step
        i-1 overlap the forward of step i; with graphs, a ring of three captured forwards with their own static
        input / output buffers, so that the copies never touch the buffers the running step is using.
        from_points: the step starts from the RAW POINT CLOUD (4.3 MB instead of 40.8 MB of voxel features):
        H2D of the points, DynamicVFE (the one host sync of the pipeline: the voxel count, as in the reference's
        torch.unique), backbone, HeightCompression (dense BEV, stays on the device for the 2-D backbone), D2H of
        the sparse output rows."""
        args, model, device, host = self.args, self.model, self.device, self.host
        if from_points:
            vfe, bev, points = self.points_pipeline()
            pts_dev = [torch.empty_like(points[0], device=device) for _ in range(3)]
        out_feat = [torch.empty((N_VOXELS, 64), dtype=torch.float32).pin_memory() for _ in range(2)]
        out_idx = [torch.empty((N_VOXELS, 4), dtype=torch.int32).pin_memory() for _ in range(2)]
        s_in, s_comp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        gpu_spans, slots = [], []
        if graph:
            for r in range(3):
                fb, cb = self.dev[r][0].clone(), self.dev_idx[r].clone()
                slots.append({"f": fb, "c": cb, "done": None, "drained": None,
                              "g": model.capture({"voxel_features": fb, "voxel_coords": cb, "batch_size": 1})})

        def run(steps, stamps=None):
            staged, keep, rows = {}, [], 0
            gpu_spans.clear()
            for sl in slots:
                sl["done"] = sl["drained"] = None

            def stage(i):
                f, c = host[i % POOL]
                with torch.cuda.stream(s_in):
                    if from_points:
                        sl = slots[i % 3] if slots else None
                        if sl is not None and sl["done"] is not None:
                            s_in.wait_event(sl["done"])
                        pts_dev[i % 3].copy_(points[i % POOL], non_blocking=True)
                        fd, cd = pts_dev[i % 3], None
                    elif slots:
                        sl = slots[i % 3]
                        if sl["done"] is not None:
                            s_in.wait_event(sl["done"])      # step i-3 has consumed the slot's inputs
                        sl["f"].copy_(f, non_blocking=True)
                        sl["c"].copy_(c, non_blocking=True)
                        fd, cd = sl["f"], sl["c"]
                    else:
                        fd = f.to(device, non_blocking=True)
                        cd = c.to(device, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(s_in)
                staged[i] = (fd, cd, ev)

            def drain(i, sp, done):
                # result of step i to the host: row count (already on its way), then the rows
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    n = sp.features.shape[0]
                    out_feat[i % 2][:n].copy_(sp.features, non_blocking=True)
                    out_idx[i % 2][:n].copy_(sp.indices, non_blocking=True)
                    if slots:
                        slots[i % 3]["drained"] = torch.cuda.Event()
                        slots[i % 3]["drained"].record(s_out)
                return n

            stage(0)
            pending = None
            for i in range(steps):
                if i + 1 < steps:
                    stage(i + 1)
                fd, cd, ev = staged.pop(i)
                with torch.cuda.stream(s_comp):
                    s_comp.wait_event(ev)
                    if slots and slots[i % 3]["drained"] is not None:
                        s_comp.wait_event(slots[i % 3]["drained"])   # step i-3's rows have left the slot
                    if stamps is not None:
                        g0 = torch.cuda.Event(enable_timing=True)
                        g0.record(s_comp)
                    if from_points:
                        vox = vfe({"points": fd, "batch_size": 1})          # (host sync: the voxel count)
                        if slots and vox["voxel_features"].shape[0] == N_VOXELS:
                            slots[i % 3]["f"].copy_(vox["voxel_features"])
                            slots[i % 3]["c"].copy_(vox["voxel_coords"])
                            sp = slots[i % 3]["g"].replay()
                        else:
                            sp = model({"voxel_features": vox["voxel_features"], "voxel_coords": vox["voxel_coords"],
                                        "batch_size": 1})["encoded_spconv_tensor"]
                        keep_bev = bev({"encoded_spconv_tensor": sp, "encoded_spconv_tensor_stride": 1})["spatial_features"]
                    elif slots:
                        sp = slots[i % 3]["g"].replay()
                    else:
                        sp = model({"voxel_features": fd, "voxel_coords": cd, "batch_size": 1})["encoded_spconv_tensor"]
                    sp.prefetch_row_count()
                    done = torch.cuda.Event(enable_timing=stamps is not None)
                    done.record(s_comp)
                    if slots:
                        slots[i % 3]["done"] = done
                    if stamps is not None:
                        gpu_spans.append((g0, done))
                if pending is not None:      # step i is queued: now wait for step i-1's count and copy it out
                    rows = drain(*pending)
                pending = (i, sp, done)
                if stamps is not None:
                    stamps.append(time.perf_counter())
                keep.append((fd, cd, sp, keep_bev if from_points else None))    # keep device buffers alive until their copies are done
                if len(keep) > 4:
                    keep.pop(0)
            rows = drain(*pending)
            for st in (s_in, s_comp, s_out):
                st.synchronize()
            return rows

        run(max(args.warmup, 8))  # (long enough for the caching allocator to reach its steady state)
        samples, rows = [], 0
        for k in range(passes):
            self.barrier()
            stamps = [time.perf_counter()]
            rows = run(args.steps, stamps)
            self.barrier()
            stamps.append(time.perf_counter())
            samples.append(stamps[-1] - stamps[0])
            busy = [a.elapsed_time(b) for a, b in gpu_spans]
            idle = [gpu_spans[j][1].elapsed_time(gpu_spans[j + 1][0]) for j in range(len(gpu_spans) - 1)]
            print(f"[bench] e2e{'_points' if from_points else ''} pass {k + 1}: {samples[-1] * 1e3:.2f} ms for {args.steps} steps; forward on the compute "
                  f"stream {statistics.mean(busy):.3f} ms/step, idle between forwards {statistics.mean(idle):.3f} ms/step",
                  file=sys.stderr)
        return sorted(samples)[len(samples) // 2], [round(v / args.steps * 1e3, 4) for v in samples], rows

    def host_link_probe(self, mb=64, reps=8):
        """What the host <-> device path of this box gives every rank WHEN ALL RANKS COPY AT ONCE (barrier-bracketed,
        pinned memory, one direction at a time, then both): names the limiter of the e2e numbers at N > 1."""
        import torch.distributed as dist
        n = mb * (1 << 20) // 4
        h_in, h_out = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
        d_in, d_out = torch.empty(n, dtype=torch.float32, device=self.device), torch.ones(n, dtype=torch.float32, device=self.device)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        res = {}
        for name in ("h2d", "d2h", "both"):
            for timed in (False, True):
                self.barrier()
                t0 = time.perf_counter()
                for _ in range(reps):
                    if name in ("h2d", "both"):
                        with torch.cuda.stream(s1):
                            d_in.copy_(h_in, non_blocking=True)
                    if name in ("d2h", "both"):
                        with torch.cuda.stream(s2):
                            h_out.copy_(d_out, non_blocking=True)
                s1.synchronize(); s2.synchronize()
                self.barrier()
                dt = time.perf_counter() - t0
            gbs = reps * n * 4 / dt / 1e9
            t = torch.tensor([gbs], dtype=torch.float64, device=self.device)
            if self.world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            res[name + "_gbs_per_rank"] = round(float(t[0]) / self.world, 2)
            res[name + "_gbs_all_ranks"] = round(float(t[0]), 2)
        res["note"] = "%d MiB pinned copies, %d back to back, all %d ranks at once; 'both' = each direction's rate while the other runs" % (mb, reps, self.world)
        return res

    def parity(self, want, modes):
        """The frame the CPU oracle just ran (seed 0, 150 k voxels), through the GPU arm in every mode -- eager and
        CUDA-graph replay: indices bit-exact, features within the mode's tolerance of max|oracle|."""
        f, c = synth_frame(0, N_VOXELS)
        f, c = torch.from_numpy(f).to(self.device), torch.from_numpy(c).to(self.device)
        scale = want.features.abs().max().item()
        res = {"frame": "seed 0, %d voxels (the cpu_baseline frame)" % N_VOXELS, "oracle_rows": int(want.indices.shape[0]),
               "modes": {}}
        with torch.no_grad():
            for m in modes:
                self.model.set_precision(m)
                g = self.model.capture({"voxel_features": f, "voxel_coords": c, "batch_size": 1})
                row = {}
                for how, sp in (("eager", self.eager_frame(f, c)), ("graph", g.replay())):
                    same = torch.equal(sp.indices.cpu(), want.indices)
                    err = (sp.features.cpu() - want.features).abs().max().item() / scale if same else None
                    row[how] = {"indices_equal": bool(same), "max_rel_err": err}
                row["tolerance"] = PARITY_TOL[m]
                row["ok"] = all(v["indices_equal"] and v["max_rel_err"] <= PARITY_TOL[m] for v in (row["eager"], row["graph"]))
                res["modes"][m] = row
                del g
        head = res["modes"][modes[0]]
        res.update({"mode": modes[0], "indices_equal": head["eager"]["indices_equal"] and head["graph"]["indices_equal"],
                    "max_rel_err": max(head["eager"]["max_rel_err"] or 0.0, head["graph"]["max_rel_err"] or 0.0),
                    "all_ok": all(r["ok"] for r in res["modes"].values())})
        return res

    def eager_frame(self, f, c):
        return self.model({"voxel_features": f, "voxel_coords": c.float(), "batch_size": 1})["encoded_spconv_tensor"]


def our_arm(args):
    rank, world, local = dist_env()
    numa = bind_to_gpu_numa_node(local)
    print("[bench] rank %d: GPU %d, NUMA node %s, %d CPUs" % (rank, local, numa, len(os.sched_getaffinity(0))),
          file=sys.stderr)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    arm = Arm(args, rank, world, local)
    others = [m for m in arm.model.PRECISIONS if m not in (args.precision, "fp32")] if args.modes else []
    head = arm.measure(args.precision, e2e_passes=3, sample_clocks=True, with_points=True)
    modes = {}
    for m in others:     # the other tensor-core modes: same frames, same K steps, graph-timed, e2e (one pass), rooflines
        r = arm.measure(m, e2e_passes=1)
        modes[m] = {k: r[k] for k in ("dtype", "value", "ms_per_step", "eager_ms_per_step", "gpu_launches", "e2e", "roofline",
                                      "kernels", "note")}
    arm.model.set_precision(args.precision)
    alt = None
    if args.launch == "graph":
        try:
            alt = arm.alternating_patterns(args.precision)
        except Exception as e:  # noqa: BLE001 -- a side measurement must not cost the line
            print("[bench] alternating-pattern measurement failed: %r" % (e,), file=sys.stderr)

    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
        "config": {"workload": WORKLOAD, "voxels_per_frame": N_VOXELS, "frames_per_step": world,
                   "sharding": "by frame, no collective",
                   "l2": "inputs rotate through a pool of %d distinct frames (326 MB > 126 MB L2)" % POOL,
                   "launch": {"graph": "one CUDA graph replay per forward (captured once per resident frame; every replay "
                                       "runs all kernels of the frame, geometry included)",
                              "eager": "kernel by kernel from Python"}[args.launch],
                   "eager_ms_per_step": round(head["eager_ms_per_step"], 4),
                   "precision": head["note"]},
        "e2e": head["e2e"], "e2e_points": head.get("e2e_points"), "gpu_launches": head["gpu_launches"], "clocks": head.get("clocks"),
        "roofline": head["roofline"], "kernels": head["kernels"], "output_rows": head["output_rows"],
        "modes": modes, "alt_patterns": alt,
        "host_link": arm.host_link_probe(),
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], want = cpu_baseline(budget_s=20.0)
            line["parity"] = arm.parity(want, [args.precision] + others)
        if world == 1 and not args.no_train_probe:
            try:        # a side measurement: it must not cost the line
                line["train_step"] = train_step_probe(torch.device("cuda", local))
            except Exception as e:  # noqa: BLE001
                line["train_step"] = {"error": repr(e)}
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def train_step_probe(device, steps=8, warmup=3):
    """SURVEY 8(f) rank 2 beside the headline: one training step (forward + backward + AdamW) of the same S0 backbone on
    150 k-voxel frames through the hand-written training kernels (csrc/train*.cu), CUDA-event timed after the headline is
    measured.  A side measurement; the records are benchmarks/train_step.py's (profiles/r02_train/)."""
    from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(s0_model_cfg(), 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).to(device).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    frames = []
    for i in range(2):
        f, c = synth_frame(100 + i, N_VOXELS)
        frames.append((torch.from_numpy(f).to(device), torch.from_numpy(c).to(device)))

    def step(i):
        f, c = frames[i % len(frames)]
        opt.zero_grad(set_to_none=True)
        sp = model({"voxel_features": f, "voxel_coords": c, "batch_size": 1})["encoded_spconv_tensor"]
        loss = (sp.dense() ** 2).mean()
        loss.backward()
        opt.step()

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, opt, frames
    return {"ms_per_step": ms, "voxels_per_s": N_VOXELS / (ms * 1e-3), "steps": steps, "warmup": warmup, "dtype": "tf32x3",
            "what": "forward + backward + AdamW, one %d-voxel frame per step, hand-written forward / backward kernels on compact "
                    "window lists (split-TF32 row-linear kernels), loss = mean(dense()^2); not part of `value`" % N_VOXELS}


# ----------------------------------------------------------------------------- CPU arm

def oracle_forward_fn(threads=None):
    from oracle import backbone as orc
    from oracle import ops as orc_ops
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    orc_ops.lib().orc_set_num_threads(threads)  # torchrun exports OMP_NUM_THREADS=1
    from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer
    cfg = s0_model_cfg()
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    state = {k: v.clone() for k, v in model.state_dict().items()}

    def run(feats, coords):
        with torch.no_grad():
            return orc.backbone_forward(state, cfg, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE), feats, coords, 1)
    return run


def cpu_baseline(budget_s=20.0):
    """The CPU oracle (a port of the reference path: the reference has no CPU implementation) on the host cores:
    all threads on full 150 k-voxel frames (passes until ~budget_s), and ONE thread on a bounded sample (a
    40 k-voxel crop of the same frame generator, same density).  Returns (record, oracle output of seed 0)."""
    cores = os.cpu_count() or 1
    run = oracle_forward_fn(cores)
    f, c = synth_frame(0, N_VOXELS)
    f, c = torch.from_numpy(f), torch.from_numpy(c)
    want = run(f, c)  # warm-up (MKL / OpenMP thread pools); also the parity reference
    times = []
    t_all = time.perf_counter()
    while time.perf_counter() - t_all < budget_s and len(times) < 5:
        t0 = time.perf_counter()
        run(f, c)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    n1 = 40000
    run1 = oracle_forward_fn(1)
    f1, c1 = synth_frame(0, n1, crop=(n1 / N_VOXELS) ** 0.5)
    f1, c1 = torch.from_numpy(f1), torch.from_numpy(c1)
    t0 = time.perf_counter()
    run1(f1, c1)
    t1 = time.perf_counter() - t0
    oracle_forward_fn(cores)
    cpu = ""
    try:
        cpu = [ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
    except Exception:  # noqa: BLE001
        pass
    return {"value": N_VOXELS / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d full-backbone passes over one 150000-voxel frame (median %.2f s/frame); "
                      "PyTorch-CPU fp32 + OpenMP C kernels (oracle/)" % (len(times), med),
            "single_thread": {"value": n1 / t1, "unit": UNIT, "cores": 1,
                              "sample": "one pass over a %d-voxel crop of the same generator (%.1f s)" % (n1, t1)},
            "cpu_model": cpu, "torch": torch.__version__}, want


def reference_arm(args):
    """The reference's CPU implementation of the path = the oracle port (the reference itself is CUDA-only), all
    host threads, same workload / metric / unit.  A step is one full 150 k-voxel frame when K + W of them fit
    in ~4 minutes on this host, otherwise a smaller crop of the same generator (stated in `sample`)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    run = oracle_forward_fn(cores)
    f, c = synth_frame(0, N_VOXELS)
    t0 = time.perf_counter()
    run(torch.from_numpy(f), torch.from_numpy(c))          # calibration pass (also warms the thread pools)
    t_frame = time.perf_counter() - t0
    n = N_VOXELS
    if t_frame * (args.steps + args.warmup) > 240.0:
        n = max(20000, int(N_VOXELS * 240.0 / (t_frame * (args.steps + args.warmup))) // 1000 * 1000)
    frames = []
    for i in range(2):
        f, c = synth_frame(i, n, crop=(n / N_VOXELS) ** 0.5)
        frames.append((torch.from_numpy(f), torch.from_numpy(c)))
    for i in range(args.warmup):
        run(*frames[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        run(*frames[i % 2])
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = ("each step = one full 150000-voxel frame" if n == N_VOXELS else
              "each step = a %d-voxel crop of the frame generator (same density): full frames take %.1f s on this host"
              % (n, t_frame))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "voxels_per_frame": N_VOXELS, "frames_per_step": world,
                   "sharding": "by frame, no collective",
                   "implementation": "CPU restatement of the reference path (oracle/: PyTorch-CPU fp32 + OpenMP C "
                                     "kernels) on the host cores -- the reference itself is CUDA-only", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_RESULT = None  # the real stdout; fd 1 itself is pointed at stderr while the benchmark runs


def emit(line):
    print(json.dumps(line), file=_RESULT or sys.stdout, flush=True)


def main():
    # ONE JSON line on stdout: libraries that write to fd 1 (NCCL's version banner, ...) go to stderr
    global _RESULT
    sys.stdout.flush()
    _RESULT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--launch", default="graph", choices=["graph", "eager"],
                    help="graph (default): one CUDA-graph replay per forward; eager: kernel by kernel from Python")
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "tf32", "tf32x3", "bf16x3", "bf16"],
                    help="headline precision mode; default = the module default (bf16x3: tensor-core kernels with split "
                         "bf16 operands, features within the fp32 bar of 1e-4).  The other tensor-core modes are measured "
                         "too and reported under `modes`")
    ap.add_argument("--no-modes", dest="modes", action="store_false", help="skip the other precision modes")
    ap.add_argument("--no-train-probe", action="store_true", help="skip the training-step side measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
