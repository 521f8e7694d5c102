"""Checkpoint IO in the reference's format, so that its released weights load unchanged.

Mirrors ``BaseModel.load_network`` / ``save_network`` (basicsr/models/base_model.py:280-306, :218-243) and the two
loads of ``MultiRefRestorationModel.__init__`` (basicsr/models/multi_ref_restoration_model.py:34-46:
``pretrain_network_feature_extractor`` -> net_extractor, ``pretrain_network_g`` -> net_g).  A checkpoint file is a
dict whose ``'params'`` (or ``'params_ema'``) entry is the state dict; keys saved from a DataParallel wrapper carry a
``'module.'`` prefix that is dropped on load.  The mirror modules have the reference's parameter names
(tests/test_abi.py checks them, e.g. all 350 keys of net_g), which is what makes this a drop-in.
"""
import torch


def read_state_dict(path, param_key='params'):
    """The state dict stored in a reference-format checkpoint, on the CPU, 'module.' prefixes removed.
    param_key=None takes the file's root dict; a missing 'params_ema' falls back to 'params' as the reference does."""
    blob = torch.load(path, map_location='cpu', weights_only=True)
    if param_key is not None:
        if param_key not in blob and 'params' in blob:
            param_key = 'params'
        blob = blob[param_key]
    return {(k[7:] if k.startswith('module.') else k): v for k, v in blob.items()}


def load_network(net, path, strict=True, param_key='params'):
    """Load `path` into `net`.  strict=True: names and shapes must match exactly (RuntimeError otherwise, from
    load_state_dict).  strict=False: unknown / missing names are tolerated and tensors whose shape differs are left
    at their current value.  Returns {'missing': [...], 'unexpected': [...], 'shape_mismatch': [...]}."""
    if isinstance(net, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel)):
        net = net.module
    state = read_state_dict(path, param_key)
    own = net.state_dict()
    report = {'missing': sorted(set(own) - set(state)), 'unexpected': sorted(set(state) - set(own)), 'shape_mismatch': []}
    if not strict:
        for k in sorted(set(own) & set(state)):
            if tuple(own[k].shape) != tuple(state[k].shape):
                report['shape_mismatch'].append(k)
                del state[k]
    net.load_state_dict(state, strict=strict)
    return report


def save_network(net, path, param_key='params'):
    """Write `net`'s state dict (CPU tensors) as {param_key: state_dict}, the reference's file layout."""
    if isinstance(net, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel)):
        net = net.module
    torch.save({param_key: {k: v.detach().cpu() for k, v in net.state_dict().items()}}, path)


def load_torchvision_vgg(module, state, container='vgg_net'):
    """ImageNet weights of a torchvision VGG (no batch-norm: `vgg19` for net_map's feature taps, `vgg16` for an extractor
    tower) into a mirror whose convolutions live under `<container>.<layer name>` in network order.  The reference gets
    these by constructing `torchvision.models.vgg19(pretrained=True)` / `vgg16(pretrained=True)` itself
    (basicsr/archs/vgg_arch.py:103-108, contras_multi_extractor_arch.py:24-25); there is no download here, so the
    caller supplies the file (`vgg19-dcbb9e9d.pth` / `vgg16-397923af.pth`) or the state dict.  torchvision names the
    convolutions `features.<index>.weight`; the i-th convolution of the file goes to the i-th convolution of the
    mirror.  Returns the list of mirror keys that were filled; raises on a shape mismatch."""
    if isinstance(state, (str, bytes)) or hasattr(state, '__fspath__'):
        state = torch.load(state, map_location='cpu', weights_only=True)
    if 'state_dict' in state and not any(k.startswith('features.') for k in state):
        state = state['state_dict']
    src = sorted({int(k.split('.')[1]) for k in state if k.startswith('features.') and k.endswith('.weight')})
    seq = getattr(module, container)
    dst = [name for name, m in seq._modules.items() if isinstance(m, torch.nn.Conv2d)]
    if len(src) < len(dst):
        raise RuntimeError('the VGG file has %d convolutions, the mirror needs %d' % (len(src), len(dst)))
    filled = []
    with torch.no_grad():
        for idx, name in zip(src, dst):
            conv = seq._modules[name]
            for part in ('weight', 'bias'):
                t = state['features.%d.%s' % (idx, part)]
                own = getattr(conv, part)
                if tuple(t.shape) != tuple(own.shape):
                    raise RuntimeError('%s.%s: file has %s, mirror has %s' % (name, part, tuple(t.shape), tuple(own.shape)))
                own.copy_(t)
                filled.append('%s.%s.%s' % (container, name, part))
    return filled


def load_pipeline(pipeline, net_g=None, net_extractor=None, strict=True, param_key='params', vgg19=None, vgg16=None):
    """The reference's loads into an MRefSRPipeline: `pretrain_network_g` -> net_g and
    `pretrain_network_feature_extractor` -> net_extractor (multi_ref_restoration_model.py:34-46), plus what the
    reference obtains from torchvision at construction time: `vgg19` (path or state dict of torchvision's ImageNet
    vgg19) -> net_map.vgg, and `vgg16` -> both towers of net_extractor when no extractor checkpoint is given.
    Returns the per-network reports; report['random_init'] lists the networks that are STILL at their random
    initial weights afterwards (results computed with any of them are meaningless and callers should say so)."""
    out = {}
    if net_extractor is not None:
        out['net_extractor'] = load_network(pipeline.net_extractor, net_extractor, strict, param_key)
    elif vgg16 is not None:
        out['net_extractor'] = {'vgg16': [load_torchvision_vgg(t, vgg16, 'model') for t in
                                          (pipeline.net_extractor.feature_extraction_image1,
                                           pipeline.net_extractor.feature_extraction_image2)][0]}
    if vgg19 is not None:
        out['net_map'] = {'vgg19': load_torchvision_vgg(pipeline.net_map.vgg, vgg19, 'vgg_net')}
    if net_g is not None:
        out['net_g'] = load_network(pipeline.net_g, net_g, strict, param_key)
    out['random_init'] = [n for n in ('net_extractor', 'net_map', 'net_g') if n not in out]
    return out
