"""Backbone configuration.

The reference's yaml (tools/cfgs/waymo_models/mssvt.yaml) is git-ignored and absent, so only the
schema is recoverable from pcdet/models/backbones_3d/mssvt_backbone.py:411-447.
`AttrDict` gives both access styles the reference uses on its EasyDict (`param.name`,
`model_cfg.get('HASH_SIZE')`) without needing easydict.  `s0_model_cfg` is config S0 of
SURVEY.md section 8, the configuration every benchmark in BASELINE.md is quoted on.
"""


class AttrDict(dict):
    """dict with attribute access, recursively applied to nested dicts / lists."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(AttrDict._wrap(e) for e in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__


def block_cfg(channels=(64, 128, 64), num_heads=(2, 2), window_size=((3, 3, 3), (5, 5, 5)),
              cbs_pattern=1, key_num_sample=32, max_num_win1=None, max_num_win2=None,
              use_feature_interpolation=True):
    return AttrDict(name="MixedScaleSparseTransformerBlock", channels=list(channels),
                    num_heads=list(num_heads), window_size=[list(w) for w in window_size],
                    max_num_win1=max_num_win1, max_num_win2=max_num_win2, cbs_mode="odd_even",
                    cbs_pattern=cbs_pattern, key_num_sample=key_num_sample,
                    use_feature_interpolation=use_feature_interpolation)


def compress_cfg(channels=(64, 128, 64), num_heads=(4,), window_size=((1, 1, 32),),
                 max_num_win1=None):
    return AttrDict(name="MixedScaleSparseTransformerCompressBlock", channels=list(channels),
                    num_heads=list(num_heads), window_size=[list(w) for w in window_size],
                    max_num_win1=max_num_win1)


def s0_model_cfg(hash_size=400000, z_cells=32, cbs_patterns=(1, 1, 1)):
    """3 mixed-scale blocks (3^3 / 5^3 windows, 2+2 heads, 32 FPS keys per scale) followed by one
    z-compress block (window 1x1xZ, 4 heads): (468, 468, 32) -> (468, 468, 1)."""
    blocks = [block_cfg(cbs_pattern=p) for p in cbs_patterns]
    blocks.append(compress_cfg(window_size=((1, 1, z_cells),)))
    return AttrDict(NAME="MixedScaleSparseTransformer", HASH_SIZE=hash_size,
                    NUM_OUTPUT_FEATURES=64, PARAMS=blocks)
