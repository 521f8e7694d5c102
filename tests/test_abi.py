"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/mrefsr_b200.h declares;
host-side argument validation; the product package never imports the oracle."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'mrefsr_b200.h')).read()
    return sorted(set(re.findall(r'MREFSR_API[^;(]*?\b(mrefsr_\w+)\s*\(', text)))


def test_header_symbols_exported():
    from mrefsr_b200 import _lib, build
    build.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(handle, s), s
    # the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == syms


def test_abi_version_and_errors_without_gpu():
    from mrefsr_b200 import _lib
    lib = _lib.lib()
    assert lib.mrefsr_abi_version() == 1
    # argument validation happens before any CUDA call
    rc = lib.mrefsr_pre_offsets(None, 0, 0, 0, None, None, None, None)
    assert rc < 0 and b'pre_offsets' in lib.mrefsr_last_error()
    rc = lib.mrefsr_mrapa_attention_forward(None, None, None, None, None, 1, 1, 1, 1, 1, 1, None)
    assert rc < 0


def test_cpu_tensors_raise():
    import mrefsr_b200 as M
    with pytest.raises(NotImplementedError):
        M.feature_match_index(torch.randn(8, 6, 6), torch.randn(8, 6, 6))
    with pytest.raises(NotImplementedError):
        M.modulated_deform_conv(torch.randn(1, 4, 5, 5), torch.randn(1, 18, 5, 5), torch.rand(1, 9, 5, 5),
                                torch.randn(4, 4, 3, 3), None, 1, 1, 1, 1, 1)
    with pytest.raises(NotImplementedError):
        M.mrapa_attention(torch.randn(1, 4, 2, 2), torch.randn(2, 4, 2, 2), torch.randn(2, 8, 2, 2), 2)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the package, the public header or tools/ may import, link or
    execute it (bench.py may, in its cpu_baseline / --impl reference legs only; tests/ and smoke() are the checkers)."""
    for sub in ('mrefsr_b200', 'tools', 'include'):
        for dirpath, dirs, files in os.walk(os.path.join(ROOT, sub)):
            dirs[:] = [d for d in dirs if d not in ('lib', '__pycache__')]
            for f in files:
                if f.endswith(('.py', '.cu', '.cuh', '.h', '.sh')):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r'^\s*(from|import)\s+oracle\b', text, re.M), f
                    assert 'oracle/' not in text, f
    # bench.py: only inside cpu_hot_path (the CPU-baseline / reference-arm leg)
    src = open(os.path.join(ROOT, 'bench.py')).read()
    body = src[src.index('def cpu_hot_path('):src.index('def _set_omp_threads(')]
    rest = src.replace(body, '')
    assert re.search(r'^\s*import oracle\b', body, re.M)
    assert not re.search(r'^\s*(from|import)\s+oracle\b', rest, re.M)


def test_module_state_dict_keys_match_reference(golden):
    """Drop-in contract: identical parameter names (checkpoint/resume row)."""
    import mrefsr_b200 as M
    g = golden('fusion')
    ref_keys = sorted(k.split('.sd.', 1)[1] for k in g.keys() if k.startswith('t5.sd.'))
    m = M.MRAPAFusion(nf=8, ref_nf=16)
    assert sorted(m.state_dict().keys()) == ref_keys
    d = M.DynAgg(16, 16, 3, stride=1, padding=1, deform_groups=4)
    assert sorted(d.state_dict().keys()) == ['bias', 'conv_offset_mask.bias', 'conv_offset_mask.weight', 'weight']
    assert d.kernel_size == (3, 3) and d.stride == (1, 1) and d.padding == (1, 1) and d.deform_groups == 4
    p = M.ModulatedDeformConvPack(8, 8, 3, padding=1, deformable_groups=2)
    assert sorted(p.state_dict().keys()) == ['bias', 'conv_offset.bias', 'conv_offset.weight', 'weight']


def test_mmcv_shim_install():
    import sys
    from mrefsr_b200 import mmcv_ops
    saved = {k: sys.modules.get(k) for k in ('mmcv', 'mmcv.ops')}
    try:
        ops = mmcv_ops.install(force=True)
        from mmcv.ops import ModulatedDeformConv2d, modulated_deform_conv2d  # noqa: F401
        assert ops.ModulatedDeformConv2d is mmcv_ops.ModulatedDeformConv2d
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_correspondence_arch_mirrors_reference_keys():
    """CorrespondenceGenerationArch owns `vgg.vgg_net.<layer>` parameters and mean/std buffers like the reference
    (basicsr/archs/vgg_arch.py:118-139), truncated at the deepest requested layer (relu3_1 -> conv3_1)."""
    import mrefsr_b200 as M
    net = M.CorrespondenceGenerationArch(patch_size=3, stride=1, vgg_layer_list=['relu1_1', 'relu2_1', 'relu3_1'])
    keys = sorted(net.state_dict().keys())
    assert keys == sorted(['vgg.mean', 'vgg.std'] + ['vgg.vgg_net.%s.%s' % (n, p) for n in
                                                      ('conv1_1', 'conv1_2', 'conv2_1', 'conv2_2', 'conv3_1')
                                                      for p in ('weight', 'bias')])
    feats = net.vgg(torch.rand(1, 3, 16, 16))
    assert {k: tuple(v.shape) for k, v in feats.items()} == {'relu1_1': (1, 64, 16, 16), 'relu2_1': (1, 128, 8, 8),
                                                            'relu3_1': (1, 256, 4, 4)}


def test_full_model_state_dict_keys_match_reference(golden):
    """The mirrors of the three MRefSR nets carry exactly the reference's parameter names and shapes (350 keys in
    net_g), so reference checkpoints load unchanged."""
    from mrefsr_b200.models import MRefSRPipeline
    g = golden('full_model')
    m = MRefSRPipeline()
    assert sorted(m.net_extractor.state_dict().keys()) == list(g('keys_ext'))
    assert sorted(m.net_map.state_dict().keys()) == list(g('keys_map'))
    sd = m.net_g.state_dict()
    assert sorted(sd.keys()) == list(g('keys_g')) and len(sd) == 350
    assert [str(tuple(v.shape)) for k, v in sorted(sd.items())] == list(g('shapes_g'))


@pytest.mark.parametrize('shape', [(80, 160, 160), (80, 80, 80), (80, 40, 40), (5, 75, 75), (3, 150, 150), (2, 300, 300),
                                   (2, 12, 12), (3, 7, 5), (1, 1, 1), (2, 33, 17)])
def test_dcn_tile_plan_covers_every_position_once(shape):
    """Host logic of the tcgen05 DCN kernel (no GPU): the row -> (sample, oy, ox) mapping of its 256-row CTA tiles,
    2-D patches or consecutive positions, must hit every output position exactly once, in range, and patches must
    not cost more than 3 % extra tiles over the linear mapping."""
    import ctypes
    import numpy as np
    from mrefsr_b200 import _lib
    b, ho, wo = shape
    lib = _lib.lib()
    meta = (ctypes.c_int * 4)()
    assert lib.mrefsr_dcn_tile_plan(b, ho, wo, ctypes.cast(meta, ctypes.c_void_p), None, 0) == 0
    tile2d, pw, ph, tiles = list(meta)
    lin_tiles = -(-b * ho * wo // 256)
    assert tiles >= lin_tiles and tiles * 100 <= lin_tiles * 103
    if not tile2d:
        assert tiles == lin_tiles and (pw, ph) == (256, 1)
    else:
        assert pw * ph in (64, 128, 256)
    coords = np.full((tiles * 256, 3), -7, dtype=np.int32)
    assert lib.mrefsr_dcn_tile_plan(b, ho, wo, ctypes.cast(meta, ctypes.c_void_p),
                                    coords.ctypes.data_as(ctypes.c_void_p), tiles * 256) == 0
    valid = coords[:, 0] >= 0
    assert (coords[~valid] == -1).all()
    c = coords[valid]
    assert (c[:, 0] < b).all() and (c[:, 1] >= 0).all() and (c[:, 1] < ho).all() and (c[:, 2] >= 0).all() and (c[:, 2] < wo).all()
    lin = (c[:, 0].astype(np.int64) * ho + c[:, 1]) * wo + c[:, 2]
    assert len(lin) == b * ho * wo and len(np.unique(lin)) == b * ho * wo
    if tile2d:       # rows of one patch are x-fastest and stay inside one sample
        first = coords[:pw * ph]
        ok = first[first[:, 0] >= 0]
        assert (ok[:, 0] == ok[0, 0]).all() and ok[:, 1].max() - ok[:, 1].min() < ph and ok[:, 2].max() - ok[:, 2].min() < pw
    # too small a buffer is an error, not an overrun
    assert lib.mrefsr_dcn_tile_plan(b, ho, wo, ctypes.cast(meta, ctypes.c_void_p),
                                    coords.ctypes.data_as(ctypes.c_void_p), tiles * 256 - 1) != 0


def test_missing_library_fails_loudly_with_the_build_command():
    """No fallback of any kind: with the shared library absent, the first use of an op raises and names the fix."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['MREFSR_LIB'] = '/nonexistent/libmrefsr_b200.so'\n"
            "import torch, mrefsr_b200 as M\n"
            "try:\n"
            "    M._lib.lib()\n"
            "except RuntimeError as e:\n"
            "    print('RAISED', e)\n" % ROOT)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    assert 'RAISED' in r.stdout and 'python -m mrefsr_b200.build' in r.stdout and 'no CPU / PyTorch fallback' in r.stdout


def test_torch_extension_module_exports_the_reference_names():
    """The module that replaces the reference's deform_conv_ext build: importable without a GPU, the five pybind exports
    of basicsr/ops/dcn/src/deform_conv_ext.cpp:150-164, and the reference's CPU-tensor error (:124)."""
    from mrefsr_b200 import build as B
    B.build_torch_ext()
    m = B.load_torch_ext()
    for name in ('deform_conv_forward', 'deform_conv_backward_input', 'deform_conv_backward_parameters',
                 'modulated_deform_conv_forward', 'modulated_deform_conv_backward'):
        assert callable(getattr(m, name)), name
    z = torch.zeros
    with pytest.raises(RuntimeError, match='not implemented on CPU'):
        m.modulated_deform_conv_forward(z(1, 8, 4, 4), z(8, 8, 3, 3), z(8), z(0), z(1, 18, 4, 4), z(1, 9, 4, 4), z(1, 8, 4, 4),
                                        z(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, True)
    with pytest.raises(RuntimeError, match='not implemented on CPU'):
        m.deform_conv_forward(z(1, 8, 4, 4), z(8, 8, 3, 3), z(1, 18, 4, 4), z(1, 8, 4, 4), z(0), z(0), 3, 3, 1, 1, 1, 1, 1, 1,
                              1, 1, 1)
