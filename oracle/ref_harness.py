"""Imports the UNMODIFIED reference (src/nr + src/gd/networks.py) on CPU.  TEST / BENCH-BASELINE INFRASTRUCTURE ONLY:
used by tests/, tests/golden/make_golden.py and bench.py's CPU legs; nothing under graspnerf_b200/ imports it.

Where the reference comes from: $GRASPNERF_REFERENCE, else /root/reference (the authoring container), else oracle/_ref
(the verbatim copy oracle/make_ref.py makes - git-ignored, it travels to the GPU box with the gpurun snapshot, so the
GPU box's host cores can time the REAL reference next to the CUDA path).

Shims: stub `easydict` (aggregate_net.py:4 imports it, never uses
it); Tensor.cuda -> identity (init_net.py:16-17 calls .cuda() in a ctor);
Tensor.to("cuda:0") -> cpu (ibrnet.py:444 hard-codes the device of pos_encoding).
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for c in (os.environ.get('GRASPNERF_REFERENCE'), '/root/reference', os.path.join(_HERE, '_ref')):
        if c and os.path.isdir(os.path.join(c, 'src', 'nr', 'network')):
            return c
    return os.path.join(_HERE, '_ref')


REF_ROOT = _find_root()


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


class shims:
    """Context manager: the reference's hard-coded .cuda() / "cuda:0" shims are active inside, torch is restored outside."""

    def __enter__(self):
        import torch
        self._saved = (torch.Tensor.cuda, torch.Tensor.to)
        install_shims()
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda, torch.Tensor.to = self._saved
        return False
