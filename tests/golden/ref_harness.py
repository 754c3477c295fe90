"""Imports the UNMODIFIED reference (/root/reference/src/nr) on CPU.

Only usable in the authoring container (the GPU box has no /root/reference).
Used by make_golden.py to generate the committed fixtures and by the
`-m "not gpu"` oracle-vs-reference tests (skipped when the reference is absent).

Shims: stub `easydict` (aggregate_net.py:4 imports it, never uses
it); Tensor.cuda -> identity (init_net.py:16-17 calls .cuda() in a ctor);
Tensor.to("cuda:0") -> cpu (ibrnet.py:444 hard-codes the device of pos_encoding).
"""
import os
import sys
import types

REF_ROOT = os.environ.get('GRASPNERF_REFERENCE', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, 'src', 'nr', 'network'))


_loaded = {}


def install_shims():
    """(Re-)installs the three shims; idempotent (a caller may have restored torch.Tensor.cuda / .to in between)."""
    import torch
    if 'easydict' not in sys.modules:
        m = types.ModuleType('easydict')
        m.EasyDict = dict
        sys.modules['easydict'] = m
    if '_to' not in _loaded:
        _loaded['_to'] = torch.Tensor.to
    _to = _loaded['_to']

    def _to_cpu(self, *a, **k):
        a = ['cpu' if isinstance(x, str) and x.startswith('cuda') else x for x in a]
        return _to(self, *a, **k)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = _to_cpu


def load_reference():
    """Returns (cfg dict, name2network) with the shims installed."""
    install_shims()
    if 'mods' in _loaded:
        return _loaded['mods']
    import torch
    import yaml
    for p in (os.path.join(REF_ROOT, 'src'), os.path.join(REF_ROOT, 'src', 'nr')):
        if p not in sys.path:
            sys.path.insert(0, p)
    with open(os.path.join(REF_ROOT, 'src', 'nr', 'configs', 'nrvgn_sdf.yaml')) as f:
        cfg = yaml.safe_load(f)
    from network.renderer import name2network
    _loaded['mods'] = (cfg, name2network)
    return _loaded['mods']


def build_reference_net(seed=0):
    import torch
    cfg, name2network = load_reference()
    torch.manual_seed(seed)
    net = name2network[cfg['network']](cfg).eval()
    return cfg, net
