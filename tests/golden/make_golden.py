"""Generate golden input/output vectors by running the UNMODIFIED reference code.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py
Writes tests/golden/*.npz (committed).  The reference ships no golden vectors for this
path, so these pin the oracle (tests/test_oracle_golden.py) and,
through the oracle and directly, the CUDA kernels (tests/test_*_gpu.py).

Harness-side shims (reference files untouched):
  * bare ``basicsr`` package object so basicsr/__init__.py (which imports every arch /
    dataset and dies on missing deps) is skipped; stub ``basicsr.version``;
  * ``mmcv.ops`` stub: ModulatedDeformConv2d base class + modulated_deform_conv2d routed to
    torchvision.ops.deform_conv2d (same offset/mask channel layout as
    basicsr/ops/dcn/src/deform_conv_cuda_kernel.cu:600-613); mmcv itself is un-vendored
    and unpinned in the reference (requirements.txt has no mmcv line);
  * torchvision vgg constructors patched to ignore ``pretrained=True`` (no network).
"""
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    pkg = types.ModuleType('basicsr')
    pkg.__path__ = [os.path.join(REF, 'basicsr')]
    sys.modules['basicsr'] = pkg
    ver = types.ModuleType('basicsr.version')
    ver.__version__ = '1.3.5'
    ver.__gitsha__ = 'unknown'
    sys.modules['basicsr.version'] = ver

    import torchvision
    from torchvision.ops import deform_conv2d
    from torch.nn.modules.utils import _pair

    class ModulatedDeformConv2d(nn.Module):
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                     groups=1, deform_groups=1, bias=True):
            super().__init__()
            self.in_channels, self.out_channels = in_channels, out_channels
            self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
            self.padding, self.dilation = _pair(padding), _pair(dilation)
            self.groups, self.deform_groups = groups, deform_groups
            self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
            self.bias = nn.Parameter(torch.Tensor(out_channels)) if bias else None
            n = in_channels * self.kernel_size[0] * self.kernel_size[1]
            self.weight.data.uniform_(-1. / math.sqrt(n), 1. / math.sqrt(n))
            if self.bias is not None:
                self.bias.data.zero_()

    def modulated_deform_conv2d(x, offset, mask, weight, bias, stride, padding, dilation, groups, deform_groups):
        return deform_conv2d(x, offset, weight, bias, stride, padding, dilation, mask)

    mmcv = types.ModuleType('mmcv')
    ops = types.ModuleType('mmcv.ops')
    ops.ModulatedDeformConv2d = ModulatedDeformConv2d
    ops.modulated_deform_conv2d = modulated_deform_conv2d
    mmcv.ops = ops
    sys.modules['mmcv'] = mmcv
    sys.modules['mmcv.ops'] = ops

    import torchvision.models.vgg as tvgg
    for name in ('vgg16', 'vgg19'):
        orig = getattr(tvgg, name)
        setattr(tvgg, name, (lambda o: (lambda pretrained=False, **kw: o(weights=None)))(orig))
        setattr(torchvision.models, name, getattr(tvgg, name))


def load_ref_map_util():
    spec = importlib.util.spec_from_file_location('ref_map_util', os.path.join(REF, 'basicsr/archs/ref_map_util.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def unit(c, h, w, gen):
    return F.normalize(torch.randn(c, h * w, generator=gen), dim=0).view(c, h, w)


def matcher_cases():
    g = torch.Generator().manual_seed(1234)
    cases = {}
    # (a) independent unit-norm features, the configuration CorrespondenceGenerationArch uses
    cases['rand'] = dict(fi=unit(64, 16, 18, g), fr=unit(64, 16, 18, g), kw=dict(is_norm=True, norm_input=True))
    # (b) KAT: reference = input translated by (+3 rows, +2 cols): arg-max known analytically
    base = unit(64, 20, 22, g)
    cases['shift'] = dict(fi=base[:, :16, :18].contiguous(), fr=base[:, 3:19, 2:20].contiguous(),
                          kw=dict(is_norm=True, norm_input=True))
    # (c) zero-padded reference (CUFED val pads to 500^2): exact-tie plateau, first index must win
    fr = unit(64, 16, 18, g)
    fr[:, 9:, :] = 0
    fr[:, :, 11:] = 0
    fi = unit(64, 16, 18, g)
    fi[:, 12:, :] = 0
    cases['zeropad'] = dict(fi=fi, fr=fr, kw=dict(is_norm=True, norm_input=True))
    # (d) defaults of the function signature (norm_input=False), un-normalised features, h != w
    cases['raw'] = dict(fi=torch.randn(32, 11, 15, generator=g), fr=torch.randn(32, 11, 15, generator=g),
                        kw=dict(is_norm=True, norm_input=False))
    # (e) no normalisation at all
    cases['nonorm'] = dict(fi=torch.randn(32, 10, 10, generator=g), fr=torch.randn(32, 10, 10, generator=g),
                           kw=dict(is_norm=False, norm_input=False))
    # (f) different input / reference sizes and non-unit strides (general signature)
    cases['strided'] = dict(fi=torch.randn(16, 13, 12, generator=g), fr=torch.randn(16, 15, 17, generator=g),
                            kw=dict(patch_size=3, input_stride=2, ref_stride=2, is_norm=True, norm_input=True))
    # (g) production channel count
    cases['c256'] = dict(fi=unit(256, 12, 12, g), fr=unit(256, 12, 12, g), kw=dict(is_norm=True, norm_input=True))
    return cases


def gen_matcher():
    rmu = load_ref_map_util()
    out = {}
    for name, c in matcher_cases().items():
        idx, val = rmu.feature_match_index(c['fi'], c['fr'], **c['kw'])
        out[f'{name}.fi'] = c['fi'].numpy()
        out[f'{name}.fr'] = c['fr'].numpy()
        out[f'{name}.idx'] = idx.numpy()
        out[f'{name}.val'] = val.numpy()
        out[f'{name}.kw'] = np.array(repr(c['kw']))
    # sample_patches layout
    x = torch.arange(2 * 4 * 5, dtype=torch.float32).view(2, 4, 5)
    out['patches.x'] = x.numpy()
    out['patches.y'] = rmu.sample_patches(x, 3, 1).contiguous().numpy()
    np.savez_compressed(os.path.join(OUT, 'matcher.npz'), **out)
    print('matcher.npz', {k: v.shape for k, v in out.items() if k.endswith('.idx')})


def gen_correspondence():
    from basicsr.archs.corres_generation_arch import CorrespondenceGenerationArch
    torch.manual_seed(10)
    net = CorrespondenceGenerationArch(patch_size=3, stride=1, vgg_layer_list=['relu1_1', 'relu2_1', 'relu3_1'],
                                       vgg_type='vgg19').eval()
    g = torch.Generator().manual_seed(77)
    b, c, h, w = 2, 64, 10, 12
    f1 = torch.randn(b, c, h, w, generator=g)
    f2 = torch.randn(b, c, h, w, generator=g)
    img = torch.rand(b, 3, 4 * h, 4 * w, generator=g)
    with torch.no_grad():
        pre, feats = net({'dense_features1': f1, 'dense_features2': f2}, img)
    rmu = load_ref_map_util()
    idx = []
    for i in range(b):
        a = F.normalize(f1[i].reshape(c, -1), dim=0).view(c, h, w)
        r = F.normalize(f2[i].reshape(c, -1), dim=0).view(c, h, w)
        idx.append(rmu.feature_match_index(a, r, patch_size=3, input_stride=1, ref_stride=1, is_norm=True,
                                           norm_input=True)[0])
    out = dict(f1=f1.numpy(), f2=f2.numpy(), max_idx=torch.stack(idx).numpy(),
               relu3_1=pre['relu3_1'].numpy(), relu2_1=pre['relu2_1'].numpy(), relu1_1=pre['relu1_1'].numpy())
    np.savez_compressed(os.path.join(OUT, 'correspondence.npz'), **out)
    print('correspondence.npz', {k: v.shape for k, v in out.items()})


def gen_dynagg():
    """DynAgg forward + gradients through the reference module (DCN = mmcv stub -> torchvision)."""
    from basicsr.archs.ref_mrapa_restoration_arch import DynAgg
    out = {}
    for name, (b, c, h, w, dg) in {'small': (2, 32, 9, 10, 8), 'big_offsets': (1, 16, 12, 12, 4)}.items():
        torch.manual_seed(5)
        m = DynAgg(c, c, 3, stride=1, padding=1, dilation=1, deform_groups=dg, extra_offset_mask=True)
        # reference zero-inits conv_offset_mask (offsets integer, mask 0.5); randomise so bilinear paths run
        g = torch.Generator().manual_seed(99)
        m.conv_offset_mask.weight.data = torch.randn(m.conv_offset_mask.weight.shape, generator=g) * 0.05
        m.conv_offset_mask.bias.data = torch.randn(m.conv_offset_mask.bias.shape, generator=g) * 0.5
        m.bias.data = torch.randn(c, generator=g) * 0.1
        x = torch.randn(b, c, h, w, generator=g)
        feat = torch.randn(b, c, h, w, generator=g)
        scale = 3.0 if name == 'small' else float(h)   # big_offsets: many samples leave the image
        pre = torch.round(torch.randn(b, 9, h, w, 2, generator=g) * scale)
        x.requires_grad_(True)
        feat.requires_grad_(True)
        y = m([x, feat], pre)
        go = torch.randn(y.shape, generator=g)
        # also capture the tensors that reach the DCN boundary
        conv_out = m.conv_offset_mask(feat)
        o1, o2, mk = torch.chunk(conv_out, 3, dim=1)
        offset = torch.cat((o1, o2), 1)
        pr = pre.repeat([1, dg, 1, 1, 1])
        reo = torch.zeros_like(offset)
        reo[:, 0::2] = pr[..., 1]
        reo[:, 1::2] = pr[..., 0]
        offset = (offset + reo).detach().requires_grad_(True)
        mask = torch.sigmoid(mk).detach().requires_grad_(True)
        xx = x.detach().clone().requires_grad_(True)
        wgt = m.weight.detach().clone().requires_grad_(True)
        bia = m.bias.detach().clone().requires_grad_(True)
        from mmcv.ops import modulated_deform_conv2d
        y2 = modulated_deform_conv2d(xx, offset, mask, wgt, bia, (1, 1), (1, 1), (1, 1), 1, dg)
        assert torch.equal(y2, y)
        y2.backward(go)
        d = dict(x=x, feat=feat, pre=pre, com_w=m.conv_offset_mask.weight, com_b=m.conv_offset_mask.bias,
                 weight=m.weight, bias=m.bias, conv_out=conv_out, offset=offset, mask=mask, y=y, go=go,
                 gx=xx.grad, goffset=offset.grad, gmask=mask.grad, gweight=wgt.grad, gbias=bia.grad)
        for k, v in d.items():
            out[f'{name}.{k}'] = v.detach().numpy()
        out[f'{name}.dg'] = np.array(dg)
    np.savez_compressed(os.path.join(OUT, 'dynagg.npz'), **out)
    print('dynagg.npz', [k for k in out if k.endswith('.y')])


def gen_fusion():
    """MRAPAFusion forward; hooks capture the attention core's inputs/outputs
    (ref_mrapa_restoration_arch.py:321-335) without touching the reference."""
    from basicsr.archs.ref_mrapa_restoration_arch import MRAPAFusion
    out = {}
    for name, (n, nf, ref_nf, h, w, t) in {'t5': (2, 8, 16, 8, 12, 5), 't1': (1, 8, 16, 4, 4, 1),
                                           'pad': (1, 8, 8, 7, 9, 3)}.items():
        torch.manual_seed(3)
        m = MRAPAFusion(nf=nf, ref_nf=ref_nf).eval()
        g = torch.Generator().manual_seed(21)
        target = torch.randn(n, nf, h, w, generator=g)
        refs = [torch.randn(n, ref_nf, h, w, generator=g) for _ in range(t)]
        cap = {}
        hs = [m.conv_emb1.register_forward_hook(lambda mod, i, o: cap.__setitem__('emb1', o.detach())),
              m.conv_emb2.register_forward_hook(lambda mod, i, o: cap.__setitem__('emb', o.detach())),
              m.conv_ass.register_forward_hook(lambda mod, i, o: cap.__setitem__('ass', o.detach())),
              m.spatial_attn.register_forward_hook(lambda mod, i, o: cap.__setitem__('cat', i[0].detach()))]
        with torch.no_grad():
            y = m(target, refs)
        for hk in hs:
            hk.remove()
        out[f'{name}.target'] = target.numpy()
        out[f'{name}.refs'] = torch.stack(refs, 0).numpy()
        out[f'{name}.emb_t'] = (cap['emb1'] * m.scale).numpy()
        out[f'{name}.emb'] = cap['emb'].numpy()
        out[f'{name}.ass'] = cap['ass'].numpy()
        out[f'{name}.core_out'] = cap['cat'][:, nf:].contiguous().numpy()
        out[f'{name}.y'] = y.numpy()
        out[f'{name}.t'] = np.array(t)
        for k, v in m.state_dict().items():
            out[f'{name}.sd.{k}'] = v.numpy()
    np.savez_compressed(os.path.join(OUT, 'fusion.npz'), **out)
    print('fusion.npz', [k for k in out if k.endswith('.y')])


def gen_dcnv1():
    """DCNv1 (no mask, no bias): torchvision.ops.deform_conv2d implements the rule of
    basicsr/ops/dcn/src/deform_conv_cuda_kernel.cu:84-112, 190-243 (the reference's own DCNv1 kernels need a GPU)."""
    from torchvision.ops import deform_conv2d
    g = torch.Generator().manual_seed(41)
    b, c, h, w, co, dg, groups = 2, 8, 9, 10, 12, 2, 2
    x = torch.randn(b, c, h, w, generator=g, requires_grad=True)
    off = (torch.randn(b, 2 * dg * 9, h, w, generator=g) * 2.5).requires_grad_(True)
    wgt = (torch.randn(co, c // groups, 3, 3, generator=g) * 0.2).requires_grad_(True)
    y = deform_conv2d(x, off, wgt, None, 1, 1, 1, None)
    go = torch.randn(y.shape, generator=g)
    y.backward(go)
    out = dict(x=x, offset=off, weight=wgt, y=y, go=go, gx=x.grad, goffset=off.grad, gweight=wgt.grad)
    np.savez_compressed(os.path.join(OUT, 'dcnv1.npz'), dg=np.array(dg), groups=np.array(groups),
                        **{k: v.detach().numpy() for k, v in out.items()})
    print('dcnv1.npz', tuple(y.shape))


def gen_full_model(name='full_model', b=2, r=2, H=48, W=56, seed=2024, quantize=False):
    """End-to-end reference forward (extractor -> net_map per reference -> net_g), exactly as
    basicsr/models/multi_ref_restoration_model.py:284-293 wires it, with key-seeded weights (tests/util.py).
    `full_model_lmr` is the LMR shape class of BASELINE config 3 in small: feature grids 15 / 30 / 60 are not
    multiples of 4, so MRAPAFusion reflect-pads and crops (ref_mrapa_restoration_arch.py:306-311, :348), and the
    reference count is 3.  `full_model_cfg1` is BASELINE config 1 literally (one 160x160 sample, 5 references); its
    inputs are 8-bit images (k / 255) and stored as uint8 to keep the fixture small."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from tests.util import refill_parameters
    from basicsr.archs.contras_multi_extractor_arch import ContrasMultiExtractorSep
    from basicsr.archs.corres_generation_arch import CorrespondenceGenerationArch
    from basicsr.archs.ref_mrapa_restoration_arch import MRAPARestorationNet
    ext = refill_parameters(ContrasMultiExtractorSep().eval(), 1)
    nmap = refill_parameters(CorrespondenceGenerationArch(patch_size=3, stride=1,
                                                          vgg_layer_list=['relu1_1', 'relu2_1', 'relu3_1'],
                                                          vgg_type='vgg19').eval(), 2)
    netg = refill_parameters(MRAPARestorationNet(ngf=64, n_blocks=16, groups=8).eval(), 3)
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(b, 3, H, W, generator=g)
    gt = F.avg_pool2d(F.pad(gt, (2, 2, 2, 2), mode='reflect'), 5, 1)            # band-limited image
    lq = F.interpolate(gt, scale_factor=0.25, mode='bicubic', align_corners=False).clamp(0, 1)
    up = F.interpolate(lq, scale_factor=4, mode='bicubic', align_corners=False).clamp(0, 1)
    refs = []
    for k in range(r):
        img = torch.roll(gt, shifts=(4 * (k + 1), -8 * (k + 1)), dims=(2, 3)) if k == 0 else torch.rand(b, 3, H, W, generator=g)
        refs.append(img)
    if quantize:
        q = lambda t: (t * 255).round() / 255            # noqa: E731
        lq, up, refs = q(lq), q(up), [q(t) for t in refs]
    with torch.no_grad():
        feats = ext(up, refs)
        pres, rfs = [], []
        for f, ref in zip(feats, refs):
            pre, rf = nmap(f, ref)
            pres.append(pre)
            rfs.append(rf)
        sr = netg(lq, pres, rfs)
    u8 = (lambda t: (t * 255).round().to(torch.uint8).numpy()) if quantize else (lambda t: t.numpy())
    out = dict(lq=u8(lq), up=u8(up), refs=u8(torch.stack(refs, 1)), sr=sr.numpy(),
               keys_ext=np.array(sorted(ext.state_dict().keys())), keys_map=np.array(sorted(nmap.state_dict().keys())),
               keys_g=np.array(sorted(netg.state_dict().keys())),
               shapes_g=np.array([str(tuple(v.shape)) for k, v in sorted(netg.state_dict().items())]))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name + '.npz sr', tuple(sr.shape), 'mean', float(sr.mean()), 'std', float(sr.std()), 'n_keys', len(out['keys_g']))


def gen_data():
    """Data preparation and validation metrics produced by the reference's own code:
    MultiRefCUFEDSet / MultiRefMegaDepthDataset.__getitem__ (basicsr/data/multi_ref_dataset.py) on synthetic PNGs
    written to a temporary directory, tensor2img (basicsr/utils/img_util.py), calculate_psnr / calculate_ssim
    (basicsr/metrics/psnr_ssim.py) and bgr2ycbcr (basicsr/utils/color_util.py).
    Harness-side additions: ``mmcv.impad`` (mmcv is un-vendored; its documented behaviour -- constant padding at the
    bottom / right up to `shape` -- via cv2.copyMakeBorder, as mmcv implements it) and a bare ``basicsr.data``
    package object so that basicsr/data/__init__.py (which imports every dataset) is skipped; the random shuffle and
    flips of the MegaDepth dataset are recorded by seeding `random` and replaying the same draws."""
    import random
    import tempfile
    import cv2
    mmcv = sys.modules['mmcv']

    def impad(img, shape, pad_val=0):
        return cv2.copyMakeBorder(img, 0, shape[0] - img.shape[0], 0, shape[1] - img.shape[1], cv2.BORDER_CONSTANT,
                                  value=pad_val)
    mmcv.impad = impad
    dpkg = types.ModuleType('basicsr.data')
    dpkg.__path__ = [os.path.join(REF, 'basicsr', 'data')]
    sys.modules['basicsr.data'] = dpkg
    from basicsr.data.multi_ref_dataset import MultiRefCUFEDSet, MultiRefMegaDepthDataset
    from basicsr.utils.img_util import tensor2img
    from basicsr.utils.color_util import bgr2ycbcr
    from basicsr.metrics.psnr_ssim import calculate_psnr, calculate_ssim

    rng = np.random.RandomState(7)

    def smooth_image(h, w):       # band-limited colour noise, uint8 BGR
        x = rng.rand(h // 4 + 2, w // 4 + 2, 3).astype(np.float32)
        x = cv2.resize(x, (w, h), interpolation=cv2.INTER_CUBIC)
        x = x + 0.05 * rng.randn(h, w, 3).astype(np.float32)
        return (np.clip(x, 0, 1) * 255).round().astype(np.uint8)

    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        # ---- CUFED5-style validation sample: 131 x 158 input (not a multiple of 4), five references of other sizes
        imgs = [smooth_image(131, 158)] + [smooth_image(h, w) for h, w in ((120, 160), (97, 143), (160, 101), (64, 64), (200, 90))]
        for k, im in enumerate(imgs):
            cv2.imwrite(os.path.join(tmp, '000_%d.png' % k), im)
        ds = MultiRefCUFEDSet({'dataroot_in': tmp, 'dataroot_ref': tmp, 'scale': 4, 'name': 'golden'})
        # the reference pads to 500 x 500; keep the fixture small by patching nothing: 500 x 500 x 6 uint8 is fine
        item = ds[0]
        out['cufed.in_bgr'] = imgs[0]
        for k in range(5):
            out['cufed.ref%d_bgr' % k] = imgs[k + 1]
        for key in ('img_in', 'img_in_lq', 'img_in_up', 'img_ref_list', 'img_ref_lq_list', 'img_ref_up_list'):
            out['cufed.' + key] = (item[key].numpy() * 255).round().astype(np.uint8)      # exact: all are u8 / 255
        out['cufed.original_size'] = np.array(item['original_size'])

        # ---- MegaDepth / LMR-style training sample: crops of 48 around points, 5 references, seeded shuffle + flips
        big = [smooth_image(140, 150) for _ in range(6)]
        rgb = [cv2.cvtColor(b, cv2.COLOR_BGR2RGB) for b in big]
        names = ['t.png', 'h.png', 'm1.png', 'm2.png', 'l1.png', 'l2.png']
        os.makedirs(os.path.join(tmp, 'scene'))
        for n, b in zip(names, big):
            cv2.imwrite(os.path.join(tmp, 'scene', n), b)
        pts = [(70, 60), (64, 80), (50, 50), (100, 90), (75, 75), (40, 100)]
        import pandas as pd
        df = pd.DataFrame([dict(target=names[0], H=names[1], M1=names[2], M2=names[3], L1=names[4], L2=names[5],
                                p0=str(list(pts[0])), p1=str(list(pts[1])), p2=str(list(pts[2])), p3=str(list(pts[3])),
                                p4=str(list(pts[4])), p5=str(list(pts[5])), scene='scene')])
        ann = os.path.join(tmp, 'ann.csv')
        df.to_csv(ann, index=False)
        md = MultiRefMegaDepthDataset({'dataroot_in': tmp, 'dataroot_ref': tmp, 'ann_file': ann, 'scale': 4,
                                       'gt_size': 48, 'use_flip': True, 'use_rot': True})
        for seed in (1, 2, 5):
            random.seed(seed)
            item = md[0]
            random.seed(seed)                       # replay the draws: shuffle, then hflip / vflip / rot90
            order = list(range(5))
            random.shuffle(order)
            hflip, vflip, rot90 = random.random() < 0.5, random.random() < 0.5, random.random() < 0.5
            out['md%d.order' % seed] = np.array(order)
            out['md%d.flags' % seed] = np.array([hflip, vflip, rot90])
            for key in ('img_in', 'img_in_lq', 'img_in_up', 'img_ref_list', 'img_ref_lq_list', 'img_ref_up_list'):
                out['md%d.%s' % (seed, key)] = (item[key].numpy() * 255).round().astype(np.uint8)
        for k in range(6):
            out['md.img%d_rgb' % k] = rgb[k]
        out['md.points'] = np.array(pts)

    # ---- tensor2img and the metrics
    g = torch.Generator().manual_seed(11)
    sr = torch.rand(1, 3, 45, 52, generator=g) * 1.2 - 0.1          # values outside [0, 1] are clamped
    gt = (sr + 0.05 * torch.randn(1, 3, 45, 52, generator=g)).clamp(0, 1)
    sr_img, gt_img = tensor2img([sr, gt])
    out['metric.sr'], out['metric.gt'] = sr.numpy(), gt.numpy()
    out['metric.sr_img'], out['metric.gt_img'] = sr_img, gt_img
    vals = []
    for cb in (0, 4):
        vals += [calculate_psnr(sr_img, gt_img, crop_border=cb, test_y_channel=False),
                 calculate_psnr(sr_img, gt_img, crop_border=cb, test_y_channel=True),
                 calculate_ssim(sr_img, gt_img, crop_border=cb, test_y_channel=True),
                 calculate_ssim(sr_img, gt_img, crop_border=cb, test_y_channel=False)]
    out['metric.values'] = np.array(vals, dtype=np.float64)      # [cb0: psnr, psnr_y, ssim_y, ssim] + [cb4: ...]
    out['metric.y_u8'] = bgr2ycbcr(sr_img, y_only=True)
    out['metric.y_f32'] = bgr2ycbcr(sr_img.astype(np.float32) / 255., y_only=True)
    np.savez_compressed(os.path.join(OUT, 'data.npz'), **out)
    print('data.npz', len(out), 'arrays; metric values', vals[:4])


if __name__ == '__main__':
    torch.set_num_threads(8)
    install_shims()
    only = set(sys.argv[1:])      # e.g. `make_golden.py full_model_lmr` regenerates one fixture
    gens = dict(matcher=gen_matcher, correspondence=gen_correspondence, dynagg=gen_dynagg, fusion=gen_fusion,
                dcnv1=gen_dcnv1, full_model=gen_full_model, data=gen_data,
                full_model_lmr=lambda: gen_full_model('full_model_lmr', b=1, r=3, H=60, W=60, seed=2025),
                full_model_cfg1=lambda: gen_full_model('full_model_cfg1', b=1, r=5, H=160, W=160, seed=2026, quantize=True))
    for key, fn in gens.items():
        if not only or key in only:
            fn()
    print('sizes:', {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT)) if f.endswith('.npz')})
