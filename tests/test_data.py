"""Data preparation and validation metrics (mrefsr_b200/data.py) against fixtures produced
by the reference's own dataset classes and metric functions (tests/golden/make_golden.py::gen_data).  CPU only."""
import math

import numpy as np
import pytest
import torch

from mrefsr_b200 import data as D


def _u8(t):
    """A [0, 1] tensor whose values are k / 255 -> the uint8 array k (what the fixtures store)."""
    return (t.numpy() * 255).round().astype(np.uint8)


KEYS = ('img_in', 'img_in_lq', 'img_in_up', 'img_ref_list', 'img_ref_lq_list', 'img_ref_up_list')


def test_cufed5_sample_matches_reference_dataset(golden):
    g = golden('data')
    refs = [g('cufed.ref%d_bgr' % k).numpy() for k in range(5)]
    s = D.prepare_cufed5_sample(g('cufed.in_bgr').numpy(), refs, scale=4, pad_shape=(500, 500))
    assert s['padding'] is True and list(s['original_size']) == g('cufed.original_size').tolist() == [128, 156]
    for key in KEYS:
        want = g('cufed.' + key).numpy()
        assert s[key].dtype == torch.float32 and tuple(s[key].shape) == want.shape, key
        assert np.array_equal(_u8(s[key]), want), key                    # bit-exact: integer pipeline (PIL bicubic)
        assert float((s[key] - torch.from_numpy(want).float() / 255.).abs().max()) == 0.0, key
    assert tuple(s['img_in'].shape) == (3, 128, 156) and tuple(s['img_in_up'].shape) == (3, 500, 500)
    assert tuple(s['img_ref_list'].shape) == (5, 3, 500, 500) and tuple(s['img_ref_lq_list'].shape) == (5, 3, 125, 125)


@pytest.mark.parametrize('seed', [1, 2, 5])
def test_megadepth_sample_matches_reference_dataset(golden, seed):
    g = golden('data')
    imgs = [g('md.img%d_rgb' % k).numpy() for k in range(6)]
    pts = [tuple(p) for p in g('md.points').tolist()]
    order = g('md%d.order' % seed).tolist()
    hflip, vflip, rot90 = (bool(v) for v in np.asarray(g('md%d.flags' % seed)).tolist())
    s = D.prepare_megadepth_sample(imgs[0], imgs[1:], pts[0], pts[1:], gt_size=48, scale=4, order=order, hflip=hflip,
                                   vflip=vflip, rot90=rot90)
    for key in KEYS:
        want = g('md%d.%s' % (seed, key)).numpy()
        assert tuple(s[key].shape) == want.shape, key
        assert np.array_equal(_u8(s[key]), want), (seed, key)


def test_small_helpers():
    x = np.arange(7 * 10 * 3, dtype=np.uint8).reshape(7, 10, 3)
    assert D.mod_crop(x, 4).shape == (4, 8, 3) and D.mod_crop(x[:, :, 0], 4).shape == (4, 8)
    with pytest.raises(ValueError):
        D.mod_crop(x[None], 4)
    p = D.impad(x, (9, 12), 5)
    assert p.shape == (9, 12, 3) and np.array_equal(p[:7, :10], x) and (p[7:] == 5).all() and (p[:, 10:] == 5).all()
    with pytest.raises(ValueError):
        D.impad(x, (6, 12))
    t = D.img2tensor(x.astype(np.float32), bgr2rgb=True)
    assert tuple(t.shape) == (3, 7, 10) and torch.equal(t[0], torch.from_numpy(x[:, :, 2].astype(np.float32)))
    a, b, c = D.augment([x, x, x], hflip=True, vflip=False, rot90=True)
    assert a.shape == (10, 7, 3) and np.array_equal(a, x[:, ::-1].transpose(1, 0, 2))


def test_tensor2img_and_metrics_match_reference(golden):
    g = golden('data')
    sr, gt = g('metric.sr'), g('metric.gt')
    sr_img, gt_img = D.tensor2img([sr, gt])
    assert np.array_equal(sr_img, g('metric.sr_img').numpy()) and np.array_equal(gt_img, g('metric.gt_img').numpy())
    assert np.array_equal(D.tensor2img(sr[0]), sr_img)                     # 3-D input, single tensor
    assert np.array_equal(D.bgr2y(sr_img), g('metric.y_u8').numpy())
    assert np.allclose(D.bgr2y(sr_img.astype(np.float32) / 255.), g('metric.y_f32').numpy(), rtol=0, atol=1e-7)
    want = g('metric.values').numpy()
    got = []
    for cb in (0, 4):
        got += [D.calculate_psnr(sr_img, gt_img, cb), D.calculate_psnr(sr_img, gt_img, cb, test_y_channel=True),
                D.calculate_ssim(sr_img, gt_img, cb, test_y_channel=True), D.calculate_ssim(sr_img, gt_img, cb)]
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9), (got, want.tolist())
    assert D.calculate_psnr(sr_img, sr_img, 0) == float('inf')
    assert abs(D.calculate_ssim(sr_img, sr_img, 0) - 1.0) < 1e-12
    chw = sr_img.transpose(2, 0, 1)
    assert D.calculate_psnr(chw, gt_img.transpose(2, 0, 1), 0, input_order='CHW') == got[0]
    with pytest.raises(ValueError):
        D.calculate_psnr(sr_img, gt_img, 0, input_order='NCHW')
    with pytest.raises(AssertionError):
        D.calculate_psnr(sr_img, gt_img[:-1], 0)
    with pytest.raises(TypeError):
        D.tensor2img(np.zeros((3, 4, 4)))


def test_evaluate_and_validate_loop(golden):
    """The per-image bookkeeping of nondist_validation: padded SR cropped back to the original size, three metrics;
    a forward that returns the bicubic input must score exactly what the metrics say about that image."""
    g = golden('data')
    refs = [g('cufed.ref%d_bgr' % k).numpy() for k in range(5)]
    s = D.prepare_cufed5_sample(g('cufed.in_bgr').numpy(), refs)
    calls = []

    def forward(lq, up, refs_):
        calls.append((tuple(lq.shape), tuple(up.shape), tuple(refs_.shape)))
        return up
    avg, per = D.validate(forward, [s, s], crop_border=4)
    assert calls == [((1, 3, 125, 125), (1, 3, 500, 500), (1, 5, 3, 500, 500))] * 2 and len(per) == 2
    up_img = D.tensor2img(s['img_in_up'])[:128, :156]
    gt_img = D.tensor2img(s['img_in'])
    assert avg['psnr'] == pytest.approx(D.calculate_psnr(up_img, gt_img, 4))
    assert avg['psnr_y'] == pytest.approx(D.calculate_psnr(up_img, gt_img, 4, test_y_channel=True))
    assert avg['ssim_y'] == pytest.approx(D.calculate_ssim(up_img, gt_img, 4, test_y_channel=True))
    assert 15 < avg['psnr'] < 60 and 0 < avg['ssim_y'] <= 1 and math.isfinite(avg['psnr_y'])
    m = D.evaluate_sr(s['img_in'], s['img_in'])
    assert m['psnr'] == float('inf') and m['sr_img'].shape == (128, 156, 3)


def _write_png(path, rgb):
    from PIL import Image
    Image.fromarray(rgb).save(path)


def test_dataset_classes_match_reference(golden, tmp_path):
    """The file-system side: the two Dataset classes on PNGs written from the fixture images must return exactly
    what the reference's classes returned for the same files and the same `random` seed."""
    import csv
    import random
    g = golden('data')
    # ---- CUFED5
    root = tmp_path / 'cufed'
    root.mkdir()
    _write_png(str(root / '000_0.png'), g('cufed.in_bgr').numpy()[:, :, ::-1])
    for k in range(5):
        _write_png(str(root / ('000_%d.png' % (k + 1))), g('cufed.ref%d_bgr' % k).numpy()[:, :, ::-1])
    ds = D.MultiRefCUFEDSet({'dataroot_in': str(root), 'dataroot_ref': str(root), 'scale': 4, 'name': 'golden'})
    assert len(ds) == 1
    item = ds[0]
    for key in KEYS:
        assert np.array_equal(_u8(item[key]), g('cufed.' + key).numpy()), key
    assert item['lq_path'].endswith('000_multi.png') and item['padding'] and tuple(item['original_size']) == (128, 156)
    # ---- MegaDepth / LMR
    mroot = tmp_path / 'md'
    (mroot / 'scene').mkdir(parents=True)
    names = ['t.png', 'h.png', 'm1.png', 'm2.png', 'l1.png', 'l2.png']
    for k, n in enumerate(names):
        _write_png(str(mroot / 'scene' / n), g('md.img%d_rgb' % k).numpy())
    pts = g('md.points').tolist()
    ann = mroot / 'ann.csv'
    with open(ann, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['target', 'H', 'M1', 'M2', 'L1', 'L2', 'p0', 'p1', 'p2', 'p3', 'p4', 'p5', 'scene'])
        w.writerow(names + [str(list(p)) for p in pts] + ['scene'])
    md = D.MultiRefMegaDepthDataset({'dataroot_in': str(mroot), 'dataroot_ref': str(mroot), 'ann_file': str(ann), 'scale': 4,
                                     'gt_size': 48, 'use_flip': True, 'use_rot': True})
    assert len(md) == 1
    for seed in (1, 2, 5):
        random.seed(seed)
        item = md[0]
        for key in KEYS:
            assert np.array_equal(_u8(item[key]), g('md%d.%s' % (seed, key)).numpy()), (seed, key)
    # a DataLoader with the reference's validation batch size (1) collates the items
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=1)))
    assert tuple(batch['img_ref_list'].shape) == (1, 5, 3, 500, 500) and tuple(batch['img_in_lq'].shape) == (1, 3, 125, 125)
