"""CPU time to enqueue one backbone forward (no device sync inside the loop) + cProfile of the enqueue path.
usage: python benchmarks/host_overhead.py [precision]"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from mssvt_b200.config import s0_model_cfg
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
    cfg = s0_model_cfg()
    cfg["PRECISION"] = prec
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).cuda().eval()
    frames = []
    for i in range(4):
        f, c = synth_frame(100 + i, 150000)
        frames.append((torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda()))

    def step(i):
        f, c = frames[i % 4]
        return model({"voxel_features": f, "voxel_coords": c, "batch_size": 1})["encoded_spconv_tensor"]

    with torch.no_grad():
        for i in range(5):
            step(i)
        torch.cuda.synchronize()
        for rep in range(3):
            t0 = time.perf_counter()
            for i in range(20):
                step(i)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(f"enqueue {1e3 * (t1 - t0) / 20:.3f} ms/frame, with drain {1e3 * (t2 - t0) / 20:.3f} ms/frame")
        pr = cProfile.Profile()
        pr.enable()
        for i in range(20):
            step(i)
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
