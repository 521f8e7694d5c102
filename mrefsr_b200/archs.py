"""Drop-in callers of the hot path: the reference's CorrespondenceGenerationArch with its per-item Python loop
(basicsr/archs/corres_generation_arch.py:49-118) replaced by one batched matcher + pre-offset launch.

The VGG feature extractor it owns (`self.vgg`, basicsr/archs/vgg_arch.py:55-161) is plain cuDNN convolutions and
out of the hot path; a minimal mirror with the same parameter names (`vgg.vgg_net.conv1_1.weight`, ...,
buffers `vgg.mean` / `vgg.std`) is built here so that reference checkpoints load, with random initial weights
(no network access: load the ImageNet weights through `load_state_dict`), or pass your own module as `vgg`.
"""
from collections import OrderedDict

import torch
from torch import nn

from . import trunk as T
from .matcher import correspondence

_VGG19_NAMES = [
    'conv1_1', 'relu1_1', 'conv1_2', 'relu1_2', 'pool1', 'conv2_1', 'relu2_1', 'conv2_2', 'relu2_2', 'pool2',
    'conv3_1', 'relu3_1', 'conv3_2', 'relu3_2', 'conv3_3', 'relu3_3', 'conv3_4', 'relu3_4', 'pool3', 'conv4_1',
    'relu4_1', 'conv4_2', 'relu4_2', 'conv4_3', 'relu4_3', 'conv4_4', 'relu4_4', 'pool4', 'conv5_1', 'relu5_1',
    'conv5_2', 'relu5_2', 'conv5_3', 'relu5_3', 'conv5_4', 'relu5_4', 'pool5'
]
_VGG19_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512, 'M']


class VGGFeatureExtractor(nn.Module):
    """VGG19 feature taps with the reference's layer names (vgg_arch.py:55-161; vgg19 without batch-norm only)."""

    def __init__(self, layer_name_list, vgg_type='vgg19', use_input_norm=True, range_norm=False,
                 requires_grad=False):
        super().__init__()
        if vgg_type != 'vgg19':
            raise NotImplementedError('only vgg19 is mirrored (the type MRefSR uses)')
        self.layer_name_list = list(layer_name_list)
        self.use_input_norm = use_input_norm
        self.range_norm = range_norm
        max_idx = max(_VGG19_NAMES.index(v) for v in self.layer_name_list)
        layers, cin = OrderedDict(), 3
        it = iter(_VGG19_NAMES)
        for v in _VGG19_CFG:
            if v == 'M':
                layers[next(it)] = nn.MaxPool2d(kernel_size=2, stride=2)
            else:
                layers[next(it)] = nn.Conv2d(cin, v, 3, padding=1)
                layers[next(it)] = nn.ReLU(inplace=True)
                cin = v
        keep = OrderedDict((k, m) for k, m in layers.items() if _VGG19_NAMES.index(k) <= max_idx)
        self.vgg_net = nn.Sequential(keep)
        for p in self.parameters():
            p.requires_grad = requires_grad
        if not requires_grad:
            self.vgg_net.eval()
        if use_input_norm:
            self.register_buffer('mean', torch.Tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
            self.register_buffer('std', torch.Tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))

    def forward(self, x):
        if self.range_norm:
            x = (x + 1) / 2
        if self.use_input_norm:
            x = (x - self.mean) / self.std
        output = {}
        if T.fast_ok(x):    # inference: conv + [bias + ReLU] pairs fused (csrc/trunk.cu)
            T.run_sequential(self.vgg_net, x, taps=self.layer_name_list, out=output)
            return output
        for key, layer in self.vgg_net._modules.items():
            x = layer(x)
            if key in self.layer_name_list:
                output[key] = x.clone()
        return output


class CorrespondenceGenerationArch(nn.Module):
    """Same constructor and forward contract as the reference class (corres_generation_arch.py:14-118):
    forward(dense_features {'dense_features1','dense_features2'}: [B,C,h,w], img_ref_hr [B,3,H,W])
      -> (pre_offset {'relu3_1': [B,9,h,w,2], 'relu2_1': [B,9,2h,2w,2], 'relu1_1': [B,9,4h,4w,2]},
          img_ref_feat {layer: features of img_ref_hr})."""

    def __init__(self, patch_size=3, stride=1, vgg_layer_list=('relu3_1', 'relu2_1', 'relu1_1'), vgg_type='vgg19',
                 vgg=None, match_mode='auto'):
        super().__init__()
        self.patch_size = patch_size
        self.stride = stride
        self.vgg_layer_list = list(vgg_layer_list)
        self.match_mode = match_mode
        self.vgg = vgg if vgg is not None else VGGFeatureExtractor(self.vgg_layer_list, vgg_type=vgg_type)

    def forward(self, dense_features, img_ref_hr):
        pre_offset = correspondence(dense_features['dense_features1'], dense_features['dense_features2'],
                                    self.patch_size, self.stride, mode=self.match_mode)
        img_ref_feat = self.vgg(img_ref_hr)
        return pre_offset, img_ref_feat
