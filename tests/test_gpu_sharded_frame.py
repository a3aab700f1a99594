"""GPU (one device): window-set sharding of one frame (mssvt_b200/sharding.py) with the ranks emulated in
lockstep -- slab + halo per "rank", halo rows copied between the local tensors after every attention block.
Every owned output row must equal the single-GPU forward bit for bit.  (The NCCL version of the exchange runs
in benchmarks/shard_frame.py under torchrun; its index logic is the one tested here and over gloo.)"""
import pytest
import torch

from mssvt_b200.config import s0_model_cfg
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer, MixedScaleSparseTransformerCompressBlock
from mssvt_b200.sharding import SlabPlan
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,precision", [(2, "tf32"), (3, "tf32x3"), (4, "fp32")])
def test_slab_sharded_forward_equals_single_gpu_bit_for_bit(world, precision):
    cfg = s0_model_cfg()
    cfg["PRECISION"] = precision
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).cuda().eval()
    f, c = synth_frame(21, 40000, crop=0.5)
    f, c = torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda()
    with torch.no_grad():
        ref = model({"voxel_features": f, "voxel_coords": c, "batch_size": 1})["encoded_spconv_tensor"]
        ref_f, ref_i = ref.features.clone(), ref.indices.clone()
        plans = [SlabPlan(c, 3, 1, r, world) for r in range(world)]
        sps = [model._sparse_tensor(f.index_select(0, p.local_rows).contiguous(),
                                    c.index_select(0, p.local_rows).contiguous(), 1) for p in plans]

        def exchange(tensors):
            snap = [t.clone() for t in tensors]
            for r in range(world - 1):
                a, b = plans[r], plans[r + 1]
                tensors[r + 1].index_copy_(0, b.recv_left, snap[r].index_select(0, a.send_right))
                tensors[r].index_copy_(0, a.recv_right, snap[r + 1].index_select(0, b.send_left))
            # the samples' first voxels (aliased by un-masked padded FPS picks, quirk Q1) come from their owners
            alias = sum(snap[r].index_select(0, plans[r].alias_local) * plans[r].alias_mine.unsqueeze(1).float()
                        for r in range(world))
            for r in range(world):
                tensors[r].index_copy_(0, plans[r].alias_local, alias)

        for i, block in enumerate(model.backbone):
            sps = [block(sp, block_idx=i) for sp in sps]
            if isinstance(block, MixedScaleSparseTransformerCompressBlock):
                break
            exchange([sp.features for sp in sps])
            if all(getattr(sp, "_xn_ready", None) is not None for sp in sps):
                exchange([sp._xn_ready[1] for sp in sps])
        key = lambda i: (i[:, 3].long() * 1000 + i[:, 2].long())
        order = torch.argsort(key(ref_i))
        kr = key(ref_i)[order]
        total = 0
        for p, sp in zip(plans, sps):
            feats, idx = sp.features, sp.indices
            keep = (idx[:, 3] >= p.lo) & (idx[:, 3] < p.hi)
            feats, idx = feats[keep], idx[keep]
            pos = torch.searchsorted(kr, key(idx))
            assert torch.equal(kr[pos], key(idx))
            assert torch.equal(ref_f[order][pos], feats), (p.rank, (ref_f[order][pos] - feats).abs().max().item())
            total += feats.shape[0]
        assert total == ref_f.shape[0]
