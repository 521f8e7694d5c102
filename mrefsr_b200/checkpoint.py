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


def load_pipeline(pipeline, net_g=None, net_extractor=None, strict=True, param_key='params'):
    """The reference's two optional loads into an MRefSRPipeline (net_map's VGG19 comes from torchvision, not from a
    checkpoint, exactly as in the reference).  Returns the per-network reports."""
    out = {}
    if net_extractor is not None:
        out['net_extractor'] = load_network(pipeline.net_extractor, net_extractor, strict, param_key)
    if net_g is not None:
        out['net_g'] = load_network(pipeline.net_g, net_g, strict, param_key)
    return out
