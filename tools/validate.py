"""Validation run in the manner of `python basicsr/test.py -opt <yml>` for MultiRefRestorationModel
(basicsr/test.py + basicsr/models/multi_ref_restoration_model.py:316-373), on this library's pipeline:

    python tools/validate.py --dataroot datasets/CUFED5 [--net-g net_g.pth] [--net-extractor feature_extraction.pth]
                             [--crop-border 4] [--save-dir results/] [--limit N] [--synthetic N]

Builds MRefSRPipeline (channels-last trunk), loads the reference's checkpoints if given (same parameter names), walks
MultiRefCUFEDSet with batch size 1 as the reference's validation loader does, and prints per-image and average PSNR /
PSNR_Y / SSIM_Y; `--save-dir` writes the SR images.  `--synthetic N` replaces the dataset by N generated CUFED5-shaped
samples (no files needed).  Needs a GPU (the pipeline has no CPU path); its parts -- sample preparation, metrics, the
loop, checkpoint IO -- are covered by the CPU tests (tests/test_data.py, tests/test_checkpoint.py).  Exercised on a B200
with --synthetic 3 (500x500 padded samples, 161 ms / image in the network; gpurun_out r02a).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrefsr_b200 import checkpoint, data  # noqa: E402


def synthetic_samples(n, seed=0):
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        h, w = int(rng.randint(200, 500)), int(rng.randint(200, 500))
        base = rng.rand(h // 8 + 2, w // 8 + 2, 3)
        img = np.kron(base, np.ones((8, 8, 1)))[:h, :w]
        img = (np.clip(img + 0.03 * rng.randn(h, w, 3), 0, 1) * 255).round().astype(np.uint8)
        refs = [np.roll(img, (8 * (k + 1), -4 * (k + 1)), (0, 1)) if k < 2 else
                (rng.rand(h, w, 3) * 255).astype(np.uint8) for k in range(5)]
        out.append(data.prepare_cufed5_sample(img, refs))
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--dataroot')
    ap.add_argument('--net-g')
    ap.add_argument('--net-extractor')
    ap.add_argument('--vgg19', help="torchvision's ImageNet vgg19 file (vgg19-dcbb9e9d.pth) for net_map's feature taps")
    ap.add_argument('--vgg16', help="torchvision's ImageNet vgg16 file, used for the extractor when --net-extractor is not given")
    ap.add_argument('--param-key', default='params')
    ap.add_argument('--crop-border', type=int, default=4)
    ap.add_argument('--save-dir')
    ap.add_argument('--limit', type=int, default=0)
    ap.add_argument('--synthetic', type=int, default=0)
    ap.add_argument('--seed', type=int, default=10)          # manual_seed of the reference's yml files
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit('tools/validate.py: needs a CUDA device (the pipeline has no CPU path)')
    from mrefsr_b200.models import MRefSRPipeline
    torch.manual_seed(args.seed)
    net = MRefSRPipeline().eval()
    rep = checkpoint.load_pipeline(net, net_g=args.net_g, net_extractor=args.net_extractor, param_key=args.param_key,
                                   vgg19=args.vgg19, vgg16=args.vgg16)
    for name, r in rep.items():
        if name != 'random_init':
            print('loaded %s: %s' % (name, ', '.join('%d %s' % (len(v), k) for k, v in r.items())))
    if rep['random_init']:
        print('WARNING: %s left at RANDOM initial weights -- the PSNR / SSIM below are a plumbing check, not a '
              'quality measurement (the reference takes the VGG19 of net_map from torchvision: pass --vgg19)'
              % ', '.join(rep['random_init']), file=sys.stderr)
        print('WARNING: random-init networks: %s' % ', '.join(rep['random_init']))
    net = net.to('cuda').channels_last_()
    if args.synthetic:
        samples = synthetic_samples(args.synthetic)
        names = ['synthetic_%03d' % i for i in range(len(samples))]
    else:
        if not args.dataroot:
            raise SystemExit('give --dataroot or --synthetic N')
        ds = data.MultiRefCUFEDSet({'dataroot_in': args.dataroot, 'dataroot_ref': args.dataroot, 'scale': 4, 'name': 'CUFED5'})
        n = len(ds) if not args.limit else min(args.limit, len(ds))
        samples = (ds[i] for i in range(n))
        names = [os.path.splitext(os.path.basename(p))[0] for p in ds.input_list[:n]]
    tot = {'psnr': 0.0, 'psnr_y': 0.0, 'ssim_y': 0.0}
    count, t_net = 0, 0.0
    for name, s in zip(names, samples):
        lq, up, refs = (s[k][None].to('cuda') for k in ('img_in_lq', 'img_in_up', 'img_ref_list'))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sr = net(lq, up, refs)
        torch.cuda.synchronize()
        t_net += time.perf_counter() - t0
        m = data.evaluate_sr(sr, s['img_in'], s['original_size'] if s.get('padding') else None, args.crop_border)
        print('# img %s # PSNR: %.4e # PSNR_Y: %.4e # SSIM_Y: %.4e.' % (name, m['psnr'], m['psnr_y'], m['ssim_y']))
        if args.save_dir:
            from PIL import Image
            os.makedirs(args.save_dir, exist_ok=True)
            Image.fromarray(np.ascontiguousarray(m['sr_img'][:, :, ::-1])).save(os.path.join(args.save_dir, name + '.png'))
        for k in tot:
            tot[k] += m[k]
        count += 1
    if count:
        print('# Validation # PSNR: %.4e # PSNR_Y: %.4e # SSIM_Y: %.4e  (%d images, %.1f ms/image in the network)'
              % (tot['psnr'] / count, tot['psnr_y'] / count, tot['ssim_y'] / count, count, 1e3 * t_net / count))


if __name__ == '__main__':
    main()
