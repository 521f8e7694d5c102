"""Host-side data preparation and validation metrics of MRefSR: what sits either side
of the network in a real-dataset run.  CPU code, like the reference's; NumPy + Pillow only (no cv2 / mmcv needed).

Mirrors, with the same names and argument meaning where the reference has a function:

* sample preparation -- ``MultiRefCUFEDSet.__getitem__`` (basicsr/data/multi_ref_dataset.py:159-238: mod-crop,
  pad to 500x500 at the bottom / right like ``mmcv.impad``, PIL bicubic x1/4 and back, BGR->RGB CHW float tensors)
  and ``MultiRefMegaDepthDataset.__getitem__`` (:68-140: crops around annotated points, flip / transpose
  augmentation of ``basicsr/data/transforms.py:94-127``, PIL bicubic down / up).  The random choices of the
  reference (reference shuffle, flips) are explicit arguments here, so a sample is reproducible.
* ``tensor2img`` (basicsr/utils/img_util.py:37-93), ``calculate_psnr`` / ``calculate_ssim``
  (basicsr/metrics/psnr_ssim.py:11-48, 82-127, 172-200) with the BT.601 luma of ``bgr2ycbcr``
  (basicsr/utils/color_util.py:38-67), and the per-image bookkeeping of
  ``MultiRefRestorationModel.nondist_validation`` (basicsr/models/multi_ref_restoration_model.py:316-361: crop the
  padding away, PSNR on BGR, PSNR and SSIM on Y).

``mmcv.impad`` is a third-party function absent from /root/reference (mmcv is un-vendored and unpinned, SURVEY
section 8c); its published behaviour -- pad at the bottom and right up to ``shape`` with ``pad_val`` -- is restated.
Parity: tests/test_data.py against fixtures produced by the reference's own functions (tests/golden/make_golden.py).
"""
import numpy as np
import torch
from PIL import Image


# ---------------------------------------------------------------------------------------------------------
# sample preparation
def mod_crop(img, scale):
    """Drop the bottom / right remainder so that both sides are multiples of `scale` (transforms.py:6-23)."""
    if img.ndim not in (2, 3):
        raise ValueError('Wrong img ndim: %d.' % img.ndim)
    h, w = img.shape[:2]
    return img[:h - h % scale, :w - w % scale, ...].copy()


def impad(img, shape, pad_val=0):
    """Pad an HWC (or HW) image at the bottom and right up to shape = (h, w) with pad_val (mmcv.impad)."""
    h, w = int(shape[0]), int(shape[1])
    if img.shape[0] > h or img.shape[1] > w:
        raise ValueError('impad: image %s larger than the target %s' % (img.shape[:2], (h, w)))
    out = np.full((h, w) + img.shape[2:], pad_val, dtype=img.dtype)
    out[:img.shape[0], :img.shape[1], ...] = img
    return out


def bicubic_down_up(img_u8, scale):
    """PIL bicubic x1/scale and back to the original size, as the datasets build the LR input and its upsampled twin
    (multi_ref_dataset.py:103-106, 187-189).  uint8 HWC in, (lq uint8 [h/s, w/s, 3], up uint8 [h, w, 3]) out."""
    h, w = img_u8.shape[:2]
    lq = Image.fromarray(img_u8).resize((w // scale, h // scale), Image.BICUBIC)
    up = lq.resize((w, h), Image.BICUBIC)
    return np.array(lq), np.array(up)


def img2tensor(imgs, bgr2rgb=True, float32=True):
    """HWC ndarray (or list of them) -> CHW tensor(s), optionally swapping BGR to RGB (img_util.py:9-34)."""
    def one(img):
        if img.ndim == 3 and img.shape[2] == 3 and bgr2rgb:
            img = img[:, :, ::-1]
        t = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))
        return t.float() if float32 else t
    return [one(i) for i in imgs] if isinstance(imgs, list) else one(imgs)


def _unit(img_u8):
    return img_u8.astype(np.float32) / 255.


def prepare_cufed5_sample(img_in_bgr, refs_bgr, scale=4, pad_shape=(500, 500)):
    """One validation sample of CUFED5 from decoded uint8 BGR images (what cv2.imread returns): the dict of
    MultiRefCUFEDSet.__getitem__ without the file path -- 'img_in' (the mod-cropped ground truth, unpadded),
    'img_in_lq', 'img_in_up', 'img_ref_list' [R,3,H,W], 'img_ref_lq_list', 'img_ref_up_list', 'padding',
    'original_size'.  All tensors RGB, CHW, float32 in [0, 1]."""
    gt = mod_crop(img_in_bgr, scale)
    original_size = gt.shape[:2]
    img_in = impad(gt, pad_shape, 0)
    refs = [impad(r, pad_shape, 0) for r in refs_bgr]
    lq, up = bicubic_down_up(img_in, scale)
    refs_lq, refs_up = zip(*[bicubic_down_up(r, scale) for r in refs])
    t_gt, t_lq, t_up = img2tensor([_unit(gt), _unit(lq), _unit(up)], bgr2rgb=True)
    return {
        'img_in': t_gt, 'img_in_lq': t_lq, 'img_in_up': t_up,
        'img_ref_list': torch.stack(img2tensor([_unit(r) for r in refs], bgr2rgb=True)),
        'img_ref_lq_list': torch.stack(img2tensor([_unit(r) for r in refs_lq], bgr2rgb=True)),
        'img_ref_up_list': torch.stack(img2tensor([_unit(r) for r in refs_up], bgr2rgb=True)),
        'padding': True, 'original_size': (int(original_size[0]), int(original_size[1])),
    }


def augment(imgs, hflip=False, vflip=False, rot90=False):
    """The reference's augmentation with its three coin flips made explicit (transforms.py:116-127): horizontal
    flip, vertical flip, then transpose of H and W -- the same for every image of the list."""
    out = []
    for img in imgs:
        if hflip:
            img = img[:, ::-1]
        if vflip:
            img = img[::-1]
        if rot90:
            img = img.transpose(1, 0, 2)
        out.append(np.ascontiguousarray(img))
    return out


def prepare_megadepth_sample(img_in_rgb, refs_rgb, p0, p_refs, gt_size, scale=4, order=None, hflip=False, vflip=False,
                             rot90=False):
    """One training sample of the MegaDepth / LMR set from decoded uint8 RGB images: gt_size crops centred on the
    annotated points p = (x, y) (multi_ref_dataset.py:80-85), the references taken in `order` (the reference
    shuffles them, :86), the explicit augmentation, PIL bicubic down / up of the uint8-truncated crops (:95-106).
    Returns the dict of MultiRefMegaDepthDataset.__getitem__."""
    half = gt_size // 2

    def crop(img, p):
        return _unit(img)[p[1] - half:p[1] + half, p[0] - half:p[0] + half]
    img_in = crop(img_in_rgb, p0)
    refs = [crop(r, p) for r, p in zip(refs_rgb, p_refs)]
    if order is not None:
        refs = [refs[i] for i in order]
    imgs = augment([img_in] + refs, hflip, vflip, rot90)
    img_in, refs = imgs[0], imgs[1:]

    def down_up(img):       # the reference truncates (astype(uint8)), it does not round
        return bicubic_down_up((img * 255).astype(np.uint8), scale)
    lq, up = down_up(img_in)
    refs_lq, refs_up = zip(*[down_up(r) for r in refs])
    t_in, t_lq, t_up = img2tensor([img_in, _unit(lq), _unit(up)], bgr2rgb=False)
    return {
        'img_in': t_in, 'img_in_lq': t_lq, 'img_in_up': t_up,
        'img_ref_list': torch.stack(img2tensor(list(refs), bgr2rgb=False)),
        'img_ref_lq_list': torch.stack(img2tensor([_unit(r) for r in refs_lq], bgr2rgb=False)),
        'img_ref_up_list': torch.stack(img2tensor([_unit(r) for r in refs_up], bgr2rgb=False)),
    }


# ---------------------------------------------------------------------------------------------------------
# validation metrics
def tensor2img(tensor, rgb2bgr=True, min_max=(0, 1)):
    """[1,3,H,W] / [3,H,W] / [H,W] tensor(s) in RGB -> uint8 HWC image(s) in BGR: clamp, rescale to [0, 1],
    x255 and round half to even (img_util.py:37-93; mini-batch grids are not needed on this path)."""
    single = torch.is_tensor(tensor)
    if not single and not (isinstance(tensor, list) and all(torch.is_tensor(t) for t in tensor)):
        raise TypeError('tensor or list of tensors expected, got %s' % type(tensor))
    out = []
    for t in ([tensor] if single else tensor):
        t = t.squeeze(0).float().detach().cpu().clamp(*min_max)
        t = (t - min_max[0]) / (min_max[1] - min_max[0])
        if t.dim() == 3:
            img = t.numpy().transpose(1, 2, 0)
            if img.shape[2] == 1:
                img = img[:, :, 0]
            elif rgb2bgr:
                img = img[:, :, ::-1]
        elif t.dim() == 2:
            img = t.numpy()
        else:
            raise TypeError('Only 3D or 2D tensors (after removing a batch of 1); got %d dimensions' % t.dim())
        out.append((img * 255.0).round().astype(np.uint8))
    return out[0] if len(out) == 1 else out


def bgr2y(img):
    """BT.601 luma of a BGR image, `bgr2ycbcr(img, y_only=True)` (color_util.py:38-67): uint8 in -> uint8 out
    (rounded), float32 in [0, 1] -> float32 in [0, 1]."""
    if img.dtype == np.uint8:
        x = img.astype(np.float32) / 255.
    elif img.dtype == np.float32:
        x = img
    else:
        raise TypeError('The img type should be np.float32 or np.uint8, but got %s' % img.dtype)
    y = np.dot(x, [24.966, 128.553, 65.481]) + 16.0
    return y.round().astype(np.uint8) if img.dtype == np.uint8 else (y / 255.).astype(np.float32)


def _to_y_channel(img):
    """metric_util.py:32-45: [0, 255] BGR -> [0, 255] Y, float, not rounded, trailing channel axis kept."""
    x = img.astype(np.float32) / 255.
    if x.ndim == 3 and x.shape[2] == 3:
        x = bgr2y(x)[..., None]
    return x * 255.


def _prepare_pair(img, img2, crop_border, input_order, test_y_channel):
    if img.shape != img2.shape:
        raise AssertionError('Image shapes are different: %s, %s.' % (img.shape, img2.shape))
    if input_order not in ('HWC', 'CHW'):
        raise ValueError('Wrong input_order %s. Supported input_orders are "HWC" and "CHW"' % input_order)
    pair = []
    for x in (img, img2):
        if x.ndim == 2:
            x = x[..., None]
        elif input_order == 'CHW':
            x = x.transpose(1, 2, 0)
        if crop_border != 0:
            x = x[crop_border:-crop_border, crop_border:-crop_border, ...]
        if test_y_channel:
            x = _to_y_channel(x)
        pair.append(x.astype(np.float64))
    return pair


def calculate_psnr(img, img2, crop_border, input_order='HWC', test_y_channel=False):
    """PSNR of two [0, 255] images (psnr_ssim.py:11-48); inf when they are equal."""
    a, b = _prepare_pair(img, img2, crop_border, input_order, test_y_channel)
    mse = np.mean((a - b) ** 2)
    return float('inf') if mse == 0 else float(10. * np.log10(255. * 255. / mse))


def _gauss_valid(x, k):
    """'valid' correlation of a 2-D float64 image with the separable kernel k (x) k."""
    n = len(k)
    rows = sum(k[i] * x[i:x.shape[0] - n + 1 + i, :] for i in range(n))
    return sum(k[j] * rows[:, j:rows.shape[1] - n + 1 + j] for j in range(n))


def _ssim_channel(x, y):
    """SSIM of one channel with the 11x11 Gaussian window (sigma 1.5) on its valid region (psnr_ssim.py:172-200)."""
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    g = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 ** 2))
    g /= g.sum()                                   # cv2.getGaussianKernel(11, 1.5)
    mu1, mu2 = _gauss_valid(x, g), _gauss_valid(y, g)
    s1 = _gauss_valid(x * x, g) - mu1 * mu1
    s2 = _gauss_valid(y * y, g) - mu2 * mu2
    s12 = _gauss_valid(x * y, g) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2))
    return m.mean()


def calculate_ssim(img, img2, crop_border, input_order='HWC', test_y_channel=False):
    """Mean SSIM over the channels of two [0, 255] images (psnr_ssim.py:82-127)."""
    a, b = _prepare_pair(img, img2, crop_border, input_order, test_y_channel)
    return float(np.mean([_ssim_channel(a[..., c], b[..., c]) for c in range(a.shape[2])]))


def evaluate_sr(sr, gt, original_size=None, crop_border=0):
    """Per-image bookkeeping of nondist_validation (multi_ref_restoration_model.py:328-361): SR and ground-truth
    tensors -> uint8 BGR images, the SR image cropped back to `original_size` when the sample was padded, then PSNR on
    BGR, PSNR on Y and SSIM on Y.  Returns {'psnr', 'psnr_y', 'ssim_y', 'sr_img'}."""
    sr_img, gt_img = tensor2img([sr, gt])
    if original_size is not None:
        sr_img = sr_img[:original_size[0], :original_size[1]]
    return {
        'psnr': calculate_psnr(sr_img, gt_img, crop_border, test_y_channel=False),
        'psnr_y': calculate_psnr(sr_img, gt_img, crop_border, test_y_channel=True),
        'ssim_y': calculate_ssim(sr_img, gt_img, crop_border, test_y_channel=True),
        'sr_img': sr_img,
    }


def validate(forward, samples, crop_border=0, device=None):
    """The validation loop (batch size 1, as basicsr/data/__init__.py:78 builds the val loader): run
    `forward(lq [1,3,h,w], up [1,3,H,W], refs [1,R,3,H,W]) -> [1,3,H,W]` on every prepared sample and average the
    three metrics.  Returns (averages dict, per-image list)."""
    per_image = []
    for s in samples:
        lq, up, refs = s['img_in_lq'][None], s['img_in_up'][None], s['img_ref_list'][None]
        if device is not None:
            lq, up, refs = lq.to(device), up.to(device), refs.to(device)
        with torch.no_grad():
            sr = forward(lq, up, refs)
        m = evaluate_sr(sr, s['img_in'], s.get('original_size') if s.get('padding') else None, crop_border)
        per_image.append({k: m[k] for k in ('psnr', 'psnr_y', 'ssim_y')})
    n = max(1, len(per_image))
    avg = {k: sum(p[k] for p in per_image) / n for k in ('psnr', 'psnr_y', 'ssim_y')}
    return avg, per_image


# ---------------------------------------------------------------------------------------------------------
# the file-system side: the two dataset classes, same `opt` keys and returned dicts as the reference
def _read_rgb(path):
    """Decoded 8-bit RGB image (PNG / JPEG) as uint8 HWC."""
    with Image.open(path) as im:
        return np.array(im.convert('RGB'))


class MultiRefCUFEDSet(torch.utils.data.Dataset):
    """CUFED5 validation set (multi_ref_dataset.py:143-238): `<name>_0.png` inputs under opt['dataroot_in'],
    `<name>_1.png` .. `<name>_5.png` references under opt['dataroot_ref'], opt['scale'].  Every item is padded to
    500x500 and carries 'lq_path', 'padding' and 'original_size' for the validation loop."""

    def __init__(self, opt):
        import glob
        import os.path as osp
        self.opt = opt
        self.input_list = sorted(glob.glob(osp.join(opt['dataroot_in'], '*_0.png')))
        self.ref_lists = [sorted(glob.glob(osp.join(opt['dataroot_ref'], '*_%d.png' % k))) for k in range(1, 6)]

    def __len__(self):
        return len(self.input_list)

    def __getitem__(self, idx):
        img_in = _read_rgb(self.input_list[idx])[:, :, ::-1]                       # BGR, as cv2.imread decodes
        refs = [_read_rgb(lst[idx])[:, :, ::-1] for lst in self.ref_lists]
        item = prepare_cufed5_sample(img_in, refs, self.opt['scale'], (500, 500))
        item['lq_path'] = self.ref_lists[0][idx].replace('_1.png', '_multi.png')
        return item


class MultiRefMegaDepthDataset(torch.utils.data.Dataset):
    """MegaDepth / LMR training set (multi_ref_dataset.py:19-140): opt['ann_file'] is a CSV with the columns target,
    H, M1, M2, L1, L2 (file names under opt['dataroot_in']/<scene>/), p0 .. p5 (the (x, y) crop centres as Python
    literals) and scene; opt['gt_size'], opt['scale'], opt['use_flip'], opt['use_rot'].  The random draws come from
    the `random` module in the reference's order (shuffle of the five references, then one draw per enabled flip),
    so `random.seed` reproduces the reference's sample."""

    def __init__(self, opt):
        import csv
        import os.path as osp
        from ast import literal_eval
        self.opt = opt
        self.samples = []
        with open(opt['ann_file'], newline='') as f:
            for row in csv.DictReader(f):
                scene = row['scene']
                target = osp.join(opt['dataroot_in'], scene, row['target'])
                refs = [osp.join(opt['dataroot_in'], scene, row[k]) for k in ('H', 'M1', 'M2', 'L1', 'L2')]
                pts = [tuple(literal_eval(row['p%d' % k])) for k in range(6)]
                self.samples.append((target, refs, pts[0], pts[1:]))

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, index):
        import random
        in_path, ref_paths, p0, p_refs = self.samples[index]
        img_in = _read_rgb(in_path)
        refs = [_read_rgb(p) for p in ref_paths]
        order = list(range(len(refs)))
        random.shuffle(order)                                      # multi_ref_dataset.py:86
        use_flip, use_rot = self.opt['use_flip'], self.opt['use_rot']
        hflip = bool(use_flip and random.random() < 0.5)           # transforms.py:116-118 (short-circuit: no draw when off)
        vflip = bool(use_rot and random.random() < 0.5)
        rot90 = bool(use_rot and random.random() < 0.5)
        return prepare_megadepth_sample(img_in, refs, p0, p_refs, self.opt['gt_size'], self.opt['scale'], order, hflip,
                                        vflip, rot90)
