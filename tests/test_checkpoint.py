"""Checkpoint IO in the reference's file format (mrefsr_b200/checkpoint.py).  CPU only."""
import pytest
import torch

from mrefsr_b200 import checkpoint as C
from mrefsr_b200.fusion import MRAPAFusion
from tests.util import refill_parameters


def _same(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    return sorted(sa) == sorted(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)


def test_roundtrip_and_reference_file_conventions(tmp_path):
    src = refill_parameters(MRAPAFusion(nf=8, ref_nf=16), 3)
    dst = refill_parameters(MRAPAFusion(nf=8, ref_nf=16), 4)
    assert not _same(src, dst)
    p = str(tmp_path / 'net.pth')
    C.save_network(src, p)
    blob = torch.load(p, weights_only=True)
    assert list(blob) == ['params'] and sorted(blob['params']) == sorted(src.state_dict())
    rep = C.load_network(dst, p)
    assert rep == {'missing': [], 'unexpected': [], 'shape_mismatch': []} and _same(src, dst)
    # saved from a DataParallel wrapper ('module.' prefix), EMA weights requested but only 'params' present
    p2 = str(tmp_path / 'dp.pth')
    torch.save({'params': {'module.' + k: v for k, v in src.state_dict().items()}}, p2)
    dst2 = refill_parameters(MRAPAFusion(nf=8, ref_nf=16), 5)
    C.load_network(dst2, p2, param_key='params_ema')
    assert _same(src, dst2)
    # root-level state dict (param_key=None)
    p3 = str(tmp_path / 'root.pth')
    torch.save(dict(src.state_dict()), p3)
    dst3 = refill_parameters(MRAPAFusion(nf=8, ref_nf=16), 6)
    C.load_network(dst3, p3, param_key=None)
    assert _same(src, dst3)


def test_strictness(tmp_path):
    src = refill_parameters(MRAPAFusion(nf=8, ref_nf=16), 3)
    state = dict(src.state_dict())
    k_drop = sorted(state)[0]
    k_shape = sorted(state)[1]
    state.pop(k_drop)
    state['extra.weight'] = torch.zeros(2)
    state[k_shape] = torch.zeros(tuple(s + 1 for s in state[k_shape].shape))
    p = str(tmp_path / 'odd.pth')
    torch.save({'params': state}, p)
    dst = refill_parameters(MRAPAFusion(nf=8, ref_nf=16), 9)
    before = {k: v.clone() for k, v in dst.state_dict().items()}
    with pytest.raises(RuntimeError):
        C.load_network(dst, p, strict=True)
    rep = C.load_network(dst, p, strict=False)
    assert rep['missing'] == [k_drop] and rep['unexpected'] == ['extra.weight'] and rep['shape_mismatch'] == [k_shape]
    after = dst.state_dict()
    assert torch.equal(after[k_drop], before[k_drop]) and torch.equal(after[k_shape], before[k_shape])
    others = [k for k in after if k not in (k_drop, k_shape)]
    assert all(torch.equal(after[k], src.state_dict()[k]) for k in others)


def test_load_pipeline_loads_the_two_networks_the_reference_loads(tmp_path):
    from mrefsr_b200.models import MRefSRPipeline
    a, b = MRefSRPipeline(), MRefSRPipeline()
    refill_parameters(a.net_g, 1)
    refill_parameters(a.net_extractor, 2)
    refill_parameters(b.net_g, 7)
    refill_parameters(b.net_extractor, 8)
    pg, pe = str(tmp_path / 'net_g.pth'), str(tmp_path / 'feature_extractor.pth')
    C.save_network(a.net_g, pg)
    C.save_network(a.net_extractor, pe)
    rep = C.load_pipeline(b, net_g=pg, net_extractor=pe)
    assert rep.pop('random_init') == ['net_map']     # its VGG19 comes from torchvision in the reference: load_pipeline(vgg19=...)
    assert set(rep) == {'net_g', 'net_extractor'} and all(not v['missing'] and not v['unexpected'] for v in rep.values())
    assert _same(a.net_g, b.net_g) and _same(a.net_extractor, b.net_extractor)
    assert len(a.net_g.state_dict()) == 350          # the reference's net_g key count (tests/test_abi.py checks the names)


def test_torchvision_vgg_weights_map_onto_the_mirrors():
    """net_map's VGG19 and the extractor towers take torchvision's ImageNet files (what the reference builds with
    pretrained=True, vgg_arch.py:103-108 / contras_multi_extractor_arch.py:24-25): i-th convolution -> i-th convolution,
    and load_pipeline reports which networks stay random."""
    import torchvision
    from mrefsr_b200.models import MRefSRPipeline
    torch.manual_seed(0)
    tv19 = torchvision.models.vgg19(weights=None).state_dict()
    tv16 = torchvision.models.vgg16(weights=None).state_dict()
    pipe = MRefSRPipeline()
    rep = C.load_pipeline(pipe)
    assert rep['random_init'] == ['net_extractor', 'net_map', 'net_g']
    rep = C.load_pipeline(pipe, vgg19=tv19, vgg16=tv16)
    assert rep['random_init'] == ['net_g']
    sd = pipe.net_map.vgg.state_dict()
    # torchvision vgg19.features: conv1_1 = 0, conv1_2 = 2, conv2_1 = 5, conv2_2 = 7, conv3_1 = 10
    for idx, name in ((0, 'conv1_1'), (2, 'conv1_2'), (5, 'conv2_1'), (7, 'conv2_2'), (10, 'conv3_1')):
        assert torch.equal(sd['vgg_net.%s.weight' % name], tv19['features.%d.weight' % idx])
        assert torch.equal(sd['vgg_net.%s.bias' % name], tv19['features.%d.bias' % idx])
    for tower in (pipe.net_extractor.feature_extraction_image1, pipe.net_extractor.feature_extraction_image2):
        sd = tower.state_dict()
        assert torch.equal(sd['model.conv3_1.weight'], tv16['features.10.weight'])
        assert torch.equal(sd['model.conv1_1.bias'], tv16['features.0.bias'])
    with pytest.raises(RuntimeError):
        C.load_torchvision_vgg(pipe.net_map.vgg, {'features.0.weight': torch.zeros(8, 3, 3, 3), 'features.0.bias': torch.zeros(8)})
