"""End-to-end parity of the MRefSR pipeline mirror (mrefsr_b200/models.py) against the reference's own forward,
run on CPU in the build container with key-seeded weights (tests/golden/make_golden.py::gen_full_model)."""
import pytest
import torch

from mrefsr_b200.models import MRefSRPipeline
from tests.util import refill_parameters

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def pipeline():
    m = MRefSRPipeline().eval()
    refill_parameters(m.net_extractor, 1)
    refill_parameters(m.net_map, 2)
    refill_parameters(m.net_g, 3)
    return m.to(DEV)


def _err(a, b, base):
    """relative error of the network's contribution (SR minus the bilinear base the net adds it to)."""
    a, b, base = a.double().cpu(), b.double().cpu(), base.double().cpu()
    return float((a - b).norm() / (b - base).norm()), float((a - b).abs().max() / (b - base).abs().max())


@pytest.mark.parametrize('path', ['fast', 'reference_order'])
def test_full_forward_matches_reference(golden, pipeline, path):
    g = golden('full_model')
    lq, up, refs, sr_ref = (g(k).to(DEV) for k in ('lq', 'up', 'refs', 'sr'))
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the plain convolutions in exact fp32, like the CPU reference
    try:
        sr = pipeline(lq, up, refs) if path == 'fast' else pipeline.forward_reference_order(lq, up, refs)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert sr.shape == sr_ref.shape
    base = torch.nn.functional.interpolate(lq, None, 4, 'bilinear', False)
    rel_l2, rel_max = _err(sr, sr_ref, base)
    # north star: final SR images within 1e-3 relative error in fp32 (here measured on the residual the network
    # produces, which is stricter than on the image itself)
    assert rel_l2 <= 1e-3, (rel_l2, rel_max)
    assert float((sr.cpu() - sr_ref.cpu()).abs().max()) <= 1e-3


def test_fast_path_equals_reference_order(pipeline):
    g = torch.Generator().manual_seed(5)
    b, r, H, W = 1, 3, 64, 64
    lq = torch.rand(b, 3, H // 4, W // 4, generator=g).to(DEV)
    up = torch.nn.functional.interpolate(lq, scale_factor=4, mode='bicubic', align_corners=False)
    refs = torch.rand(b, r, 3, H, W, generator=g).to(DEV)
    a = pipeline(lq, up, refs)
    c = pipeline.forward_reference_order(lq, up, refs)
    base = torch.nn.functional.interpolate(lq, None, 4, 'bilinear', False)
    rel_l2, _ = _err(a, c, base)
    assert rel_l2 <= 2e-3


def test_channels_last_trunk_matches_reference(golden):
    """The channels-last trunk (cuDNN NHWC kernels, DCN reading / writing NHWC, layout-following glue kernels)
    against the same golden output of the reference."""
    import copy
    m = MRefSRPipeline().eval()
    refill_parameters(m.net_extractor, 1)
    refill_parameters(m.net_map, 2)
    refill_parameters(m.net_g, 3)
    m = copy.deepcopy(m).to(DEV).channels_last_()
    g = golden('full_model')
    lq, up, refs, sr_ref = (g(k).to(DEV) for k in ('lq', 'up', 'refs', 'sr'))
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        sr = m(lq, up, refs)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    base = torch.nn.functional.interpolate(lq, None, 4, 'bilinear', False)
    rel_l2, rel_max = _err(sr, sr_ref, base)
    assert rel_l2 <= 1e-3, (rel_l2, rel_max)
    assert float((sr.cpu() - sr_ref.cpu()).abs().max()) <= 1e-3


@pytest.mark.parametrize('path', ['fast', 'reference_order', 'channels_last'])
def test_lmr_shape_class_matches_reference(golden, pipeline, path):
    """BASELINE config 3's shape class in small (tests/golden/make_golden.py::gen_full_model('full_model_lmr')):
    60x60 HR -> feature grids 15 / 30 / 60, so MRAPAFusion reflect-pads to 16 / 32 and crops
    (ref_mrapa_restoration_arch.py:306-311, :348), the matcher sees a 13x13 origin grid, and there are 3 references."""
    import copy
    g = golden('full_model_lmr')
    lq, up, refs, sr_ref = (g(k).to(DEV) for k in ('lq', 'up', 'refs', 'sr'))
    m = copy.deepcopy(pipeline).channels_last_() if path == 'channels_last' else pipeline
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        sr = m.forward_reference_order(lq, up, refs) if path == 'reference_order' else m(lq, up, refs)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert sr.shape == sr_ref.shape == (1, 3, 60, 60)
    base = torch.nn.functional.interpolate(lq, None, 4, 'bilinear', False)
    rel_l2, rel_max = _err(sr, sr_ref, base)
    assert rel_l2 <= 1e-3, (rel_l2, rel_max)
    assert float((sr.cpu() - sr_ref.cpu()).abs().max()) <= 1e-3


def test_graphed_forward_equals_eager(golden, pipeline):
    """CUDA-graph replay (MRefSRPipeline.graphed) against the eager forward: same kernels, same order, so the SR
    images must be bit-identical -- also for new inputs copied into the static buffers after the capture."""
    import copy
    g = golden('full_model')
    lq, up, refs = (g(k).to(DEV) for k in ('lq', 'up', 'refs'))
    m = copy.deepcopy(pipeline).channels_last_()
    eager = m(lq, up, refs).clone()
    runner = m.graphed(lq, up, refs)
    assert torch.equal(runner(lq, up, refs), eager)
    gen = torch.Generator().manual_seed(77)
    lq2 = torch.rand(lq.shape, generator=gen)
    up2 = torch.nn.functional.interpolate(lq2, scale_factor=4, mode='bicubic', align_corners=False).clamp(0, 1)
    refs2 = torch.rand(refs.shape, generator=gen)
    host_out = torch.empty(eager.shape).pin_memory()
    runner(lq2.pin_memory(), up2.pin_memory(), refs2.pin_memory(), out=host_out)      # host tensors in, host tensor out
    torch.cuda.synchronize()
    eager2 = m(lq2.to(DEV), up2.to(DEV), refs2.to(DEV))
    assert torch.equal(host_out, eager2.cpu())


def test_config1_literal_sample_matches_reference(golden, pipeline):
    """BASELINE config 1: one CUFED5-shaped sample, 160x160 HR, 5 references, the reference's fp32 CPU forward
    (tests/golden/make_golden.py::gen_full_model('full_model_cfg1'); inputs are 8-bit images stored as uint8)."""
    import copy
    g = golden('full_model_cfg1')
    lq, up, refs = (g(k).float().div(255).to(DEV) for k in ('lq', 'up', 'refs'))
    sr_ref = g('sr').to(DEV)
    m = copy.deepcopy(pipeline).channels_last_()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        sr = m(lq, up, refs)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert sr.shape == sr_ref.shape == (1, 3, 160, 160)
    base = torch.nn.functional.interpolate(lq, None, 4, 'bilinear', False)
    rel_l2, rel_max = _err(sr, sr_ref, base)
    assert rel_l2 <= 1e-3, (rel_l2, rel_max)
    assert float((sr.cpu() - sr_ref.cpu()).abs().max()) <= 1e-3


def test_training_forward_with_batched_references_equals_the_loop():
    """MRAPARestorationNet.forward (the training path, autograd on): the references of a scale run as one batch
    (DynamicAggregationRestoration._forward_refs_batched: split offset_conv1, one DynAgg node over B*R samples, fusion head
    on the stacked tensor) against the reference's Python loop over references (ref_mrapa_restoration_arch.py:213-259):
    same SR image, same gradients."""
    from mrefsr_b200.models import MRAPARestorationNet
    g = torch.Generator().manual_seed(9)
    b, r, h = 2, 3, 12
    net = MRAPARestorationNet(ngf=64, n_blocks=1, groups=8)
    refill_parameters(net, 4)
    net = net.to(DEV).train()
    lq = torch.rand(b, 3, h, h, generator=g).to(DEV)
    feats = [{'relu3_1': torch.randn(b, 256, h, h, generator=g).to(DEV),
              'relu2_1': torch.randn(b, 128, 2 * h, 2 * h, generator=g).to(DEV),
              'relu1_1': torch.randn(b, 64, 4 * h, 4 * h, generator=g).to(DEV)} for _ in range(r)]
    pres = [{'relu3_1': (torch.randn(b, 9, h, h, 2, generator=g) * 2).to(DEV),
             'relu2_1': (torch.randn(b, 9, 2 * h, 2 * h, 2, generator=g) * 4).to(DEV),
             'relu1_1': (torch.randn(b, 9, 4 * h, 4 * h, 2, generator=g) * 8).to(DEV)} for _ in range(r)]
    gt = torch.rand(b, 3, 4 * h, 4 * h, generator=g).to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    res = {}
    try:
        for batched in (False, True):
            net.dyn_agg_restore.batch_refs = batched
            net.zero_grad(set_to_none=True)
            out = net(lq, pres, feats)
            torch.nn.functional.l1_loss(out, gt).backward()
            res[batched] = (out.detach(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None})
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert float((res[True][0] - res[False][0]).abs().max()) <= 1e-4
    assert res[True][1].keys() == res[False][1].keys() and len(res[True][1]) > 50
    gmax = max(float(v.double().norm()) for v in res[False][1].values())
    # the same tensors handed over already stacked over the references (what MRefSRPipeline.correspondences returns)
    keys = ('relu3_1', 'relu2_1', 'relu1_1')
    feats_s = {k: torch.stack([f[k] for f in feats], 1).flatten(0, 1) for k in keys}
    pres_s = {k: torch.stack([p[k] for p in pres], 1).flatten(0, 1) for k in keys}
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        net.zero_grad(set_to_none=True)
        out_s = net(lq, pres_s, feats_s, n_refs=r)
        torch.nn.functional.l1_loss(out_s, gt).backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert float((out_s.detach() - res[True][0]).abs().max()) <= 1e-6
    for k, p in net.named_parameters():
        if p.grad is not None:      # same code path from here on; cuDNN's weight-gradient kernels are not run-to-run exact
            a, c = p.grad.double(), res[True][1][k].double()
            assert float((a - c).norm()) <= 1e-4 * float(c.norm()) + 1e-6 * gmax, k
    gmax = max(float(v.double().norm()) for v in res[False][1].values())
    for k in res[False][1]:
        a, c = res[True][1][k].double(), res[False][1][k].double()
        # TF32 DCN / GEMMs and a different batch chunking: relative to the parameter's own gradient, plus a floor for
        # the few parameters whose gradient is a sum of cancelling terms (PReLU slopes)
        assert float((a - c).norm()) <= 5e-3 * float(c.norm()) + 1e-4 * gmax, k


def test_forward_ragged_eager_and_graph_replay(pipeline):
    """BASELINE config 3 shape class: images with different reference counts.  forward_ragged buckets them by
    (R, H, W); with graphs=True each bucket replays a captured CUDA graph.  Both must return, in input order, what a
    plain forward on each image alone returns."""
    g = torch.Generator().manual_seed(12)
    H = 48
    samples = []
    for r in (2, 3, 2, 1):
        lq = torch.rand(3, H // 4, H // 4, generator=g)
        up = torch.nn.functional.interpolate(lq[None], scale_factor=4, mode='bicubic', align_corners=False)[0]
        samples.append((lq.to(DEV), up.to(DEV), torch.rand(r, 3, H, H, generator=g).to(DEV)))
    alone = [pipeline(lq[None], up[None], refs[None])[0] for lq, up, refs in samples]
    eager = pipeline.forward_ragged(samples)
    graphed = pipeline.forward_ragged(samples, graphs=True)
    again = pipeline.forward_ragged(samples, graphs=True)          # second call: replay only
    pipeline.clear_graphs()
    for a, e, gr, ag in zip(alone, eager, graphed, again):
        assert e.shape == a.shape == gr.shape
        # batch 1 vs batch 2 of the same bucket may pick different cuDNN algorithms
        assert float((e - a).abs().max()) <= 2e-3
        assert float((gr - e).abs().max()) <= 1e-5 and torch.equal(gr, ag)


def test_correspondences_batched_equals_the_per_reference_loop(pipeline):
    """MRefSRPipeline.correspondences (one extractor / matcher / VGG pass over all references) against the reference's
    per-reference loop over net_extractor / net_map: same pre-offsets (integers), same reference features."""
    g = torch.Generator().manual_seed(21)
    b, r, H = 2, 3, 48
    lq = torch.rand(b, 3, H // 4, H // 4, generator=g).to(DEV)
    up = torch.nn.functional.interpolate(lq, scale_factor=4, mode='bicubic', align_corners=False)
    refs = torch.rand(b, r, 3, H, H, generator=g).to(DEV)
    pres, feats, n = pipeline.correspondences(up, refs)
    assert n == r
    with torch.no_grad():
        ref_list = list(refs.unbind(1))
        fl = pipeline.net_extractor(up, ref_list)
        for k_ref, (f, ref) in enumerate(zip(fl, ref_list)):
            pre, rf = pipeline.net_map(f, ref)
            for key in ('relu3_1', 'relu2_1', 'relu1_1'):
                a = pres[key].unflatten(0, (b, r))[:, k_ref]
                assert a.shape == pre[key].shape
                # batch 6 vs batch 2 through cuDNN may flip a near-tie of the arg-max: allow a handful of positions
                assert float((a != pre[key]).float().mean()) <= 2e-3, key
                fa = feats[key].unflatten(0, (b, r))[:, k_ref]
                assert float((fa - rf[key]).abs().max()) <= 1e-3 * float(rf[key].abs().max()), key
