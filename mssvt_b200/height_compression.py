"""BEV consumer of the backbone output: `SparseTensor.dense()` (mssvt_dense_scatter) -> (B, C*D, H, W).

Mirror of pcdet/models/backbones_2d/map_to_bev/height_compression.py:5-51 (same constructor, config
keys, state-dict names `compress_layers.{i}` and batch_dict keys), so that a detector built on the
reference's `MAP_TO_BEV` registry can take the module from here.  The scatter is ours; the optional
3x3 Conv2d + BatchNorm2d + ReLU stack after it is plain torch.nn (library territory).
"""
import torch
import torch.nn as nn


class HeightCompression(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = model_cfg.NUM_BEV_FEATURES
        self.compress_layer_nums = model_cfg.get('COMPRESS_LAYER_NUMS', 3)
        self.layer_strides = model_cfg.get('LAYER_STRIDES', [1, 1, 1])
        self.layer_dialations = model_cfg.get('LAYER_DIALATIONS', [1, 1, 2])
        self.layer_paddings = model_cfg.get('LAYER_PADDINGS', [1, 1, 2])
        self.compress_layers = None
        if self.compress_layer_nums:
            layers = []
            for i in range(self.compress_layer_nums):
                layers += [nn.Conv2d(self.num_bev_features, self.num_bev_features, kernel_size=3,
                                     stride=self.layer_strides[i], padding=self.layer_paddings[i],
                                     dilation=self.layer_dialations[i], bias=False),
                           nn.BatchNorm2d(self.num_bev_features), nn.ReLU(inplace=True)]
            self.compress_layers = nn.ModuleList(layers)
        self.use_amp = model_cfg.get('AMP', False)

    def forward(self, batch_dict):
        with torch.autocast("cuda", enabled=bool(self.use_amp)):
            dense = batch_dict['encoded_spconv_tensor'].dense()      # (B, C, D, H, W), device-side row count
            B, C, D, H, W = dense.shape
            spatial_features = dense.view(B, C * D, H, W)
            if self.compress_layers is not None:
                for layer in self.compress_layers:
                    spatial_features = layer(spatial_features)
        batch_dict['spatial_features'] = spatial_features.float()
        batch_dict['spatial_features_stride'] = batch_dict['encoded_spconv_tensor_stride']
        return batch_dict
