"""Packs the reference's state_dict tensors into the fp32 blob the K2 kernels read.

The blob layout (entry names, offsets, padded shapes) is owned by the native library and enumerated through the C ABI
(gn_weight_entry); this module only knows which reference tensor feeds which entry and how to permute it.  Key names
are the reference's, so a `model_best.pth` packs unchanged.
"""
import numpy as np
import torch

from . import _lib

# record channel order of the 35-wide "rgb_feat": [img_feats 32 | rgb 3]; reference order is [rgb 3 | img_feats 32]
# (ibrnet.py:458-459, aggregate_net.py:67)
PERM35 = list(range(3, 35)) + [0, 1, 2]


def _np(t):
    return t.detach().to('cpu', torch.float32).numpy()


def _entries(sd, agg_prefix, dd_prefix):
    A = agg_prefix + 'agg_impl.'
    e = {}

    def lin(dst_w, dst_b, key, vector=False):
        w = _np(sd[key + '.weight'])
        e[dst_w] = w if vector else w.T
        if dst_b is not None:
            e[dst_b] = _np(sd[key + '.bias'])[None, :]
    w = _np(sd[A + 'ray_dir_fc.0.weight'])                      # [16,4]
    e['rd.w0'], e['rd.b0'] = w.T, _np(sd[A + 'ray_dir_fc.0.bias'])[None, :]
    e['rd.w1'] = _np(sd[A + 'ray_dir_fc.2.weight'])[PERM35].T   # [16,35], columns in record order
    e['rd.b1'] = _np(sd[A + 'ray_dir_fc.2.bias'])[PERM35][None, :]
    for short, name in (('mean', 'mean_decoder'), ('var', 'var_decoder'), ('aw', 'aw_decoder')):
        lin(f'dd.{short}.w0', f'dd.{short}.b0', f'{dd_prefix}{name}.0')
        lin(f'dd.{short}.w2', f'dd.{short}.b2', f'{dd_prefix}{name}.2')
        lin(f'dd.{short}.w4', f'dd.{short}.b4', f'{dd_prefix}{name}.4')
    lin('pe.w0', 'pe.b0', agg_prefix + 'prob_embed.0')
    lin('pe.w2', 'pe.b2', agg_prefix + 'prob_embed.2')
    lin('nf.w0', 'nf.b0', A + 'neuray_fc.0')
    lin('nf.w2', 'nf.b2', A + 'neuray_fc.2', vector=True)
    w = _np(sd[A + 'base_fc.0.weight'])                       # [64, 207]
    wg = np.zeros((144, 64), np.float32)
    for q in range(4):                                         # mean0 | var0 | mean1 | var1
        wg[q * 36:q * 36 + 35] = w[:, [q * 35 + c for c in PERM35]].T
    wf = np.zeros((36, 64), np.float32)
    wf[:35] = w[:, [140 + c for c in PERM35]].T
    e['bf.wg'], e['bf.wf'], e['bf.wp'] = wg, wf, w[:, 175:207].T
    e['bf.b0'] = _np(sd[A + 'base_fc.0.bias'])[None, :]
    lin('bf.w2', 'bf.b2', A + 'base_fc.2')
    lin('vf.w0', 'vf.b0', A + 'vis_fc.0')
    lin('vf.w2', 'vf.b2', A + 'vis_fc.2')
    lin('v2.w0', 'v2.b0', A + 'vis_fc2.0')
    lin('v2.w2', 'v2.b2', A + 'vis_fc2.2', vector=True)
    lin('rf.w0', 'rf.b0', A + 'rgb_fc.0')
    lin('rf.w2', 'rf.b2', A + 'rgb_fc.2')
    lin('rf.w4', 'rf.b4', A + 'rgb_fc.4', vector=True)
    lin('gf.w0', 'gf.b0', A + 'geometry_fc.0')
    lin('gf.w2', 'gf.b2', A + 'geometry_fc.2')
    for dst, src in (('at.wq', 'w_qs'), ('at.wk', 'w_ks'), ('at.wv', 'w_vs'), ('at.fc', 'fc')):
        e[dst] = _np(sd[f'{A}ray_attention.{src}.weight']).T
    e['at.ln_w'] = _np(sd[A + 'ray_attention.layer_norm.weight'])[None, :]
    e['at.ln_b'] = _np(sd[A + 'ray_attention.layer_norm.bias'])[None, :]
    lin('og.w0', 'og.b0', A + 'out_geometry_fc.0')
    lin('og.w1', 'og.b1', A + 'out_geometry_fc.1', vector=True)
    # fusions through the activation-free prob_embed.2 (fp64 products, rounded once to fp32)
    w2k, b2 = e['pe.w2'].astype(np.float64), e['pe.b2'].astype(np.float64)              # [32 in][32 out], [1,32]
    e['nfc.w0'] = (w2k @ e['nf.w0'].astype(np.float64)).astype(np.float32)
    e['nfc.b0'] = (b2 @ e['nf.w0'].astype(np.float64) + e['nf.b0']).astype(np.float32)
    e['bf.wpc'] = (w2k @ e['bf.wp'].astype(np.float64)).astype(np.float32)
    e['bf.b0c'] = (b2 @ e['bf.wp'].astype(np.float64) + e['bf.b0']).astype(np.float32)
    return e


def pack_blob(sd, agg_prefix='agg_net.', dd_prefix='dist_decoder.'):
    """-> np.float32 [gn_weight_blob_floats()]"""
    lib = _lib.load()
    table = _lib.weight_table()
    ent = _entries(sd, agg_prefix, dd_prefix)
    blob = np.zeros(lib.gn_weight_blob_floats(), np.float32)
    seen = set()
    for name, off, rows, cols, cp in table:
        a = ent[name]
        a = a.reshape(rows, cols) if a.ndim == 1 else a
        if a.shape[0] < rows:                                   # row-padded entries keep their trailing zero rows
            a = np.concatenate([a, np.zeros((rows - a.shape[0], a.shape[1]), np.float32)], 0)
        assert a.shape == (rows, cols), (name, a.shape, (rows, cols))
        view = blob[off:off + rows * cp].reshape(rows, cp)
        view[:, :cols] = a
        seen.add(name)
    assert seen == set(ent), set(ent) ^ seen
    return blob


class DevicePacker:
    """pack_blob on the device, without host round trips (training re-packs the weights every step).

    The raw entries of the blob are a fixed gather of the parameters (transposes, the [rgb|img] permutation, zero padding):
    the gather index is built ONCE by running `pack_blob` on a state_dict of element ids.  The four host-fused composites
    (nfc.w0, nfc.b0, bf.wpc, bf.b0c: products through the activation-free prob_embed.2) are evaluated with fp64 torch
    matmuls on the device and rounded once, exactly like `_entries` does with numpy."""

    def __init__(self, sd, agg_prefix, dd_prefix, device):
        self.agg_prefix, self.dd_prefix, self.device = agg_prefix, dd_prefix, torch.device(device)
        self.keys = sorted(k for k in sd if (k.startswith(agg_prefix) or k.startswith(dd_prefix)) and not k.endswith('deviation_network.variance'))
        ids, n = {}, 1                                     # id 0 = "zero" (padding)
        for k in self.keys:
            m = sd[k].numel()
            ids[k] = torch.arange(n, n + m, dtype=torch.float32).reshape(sd[k].shape)      # exact in fp32 (n << 2^24)
            n += m
        lab = pack_blob(ids, agg_prefix, dd_prefix)
        table = {name: (off, rows, cols, cp) for name, off, rows, cols, cp in _lib.weight_table()}
        self.fused = {name: table[name] for name in ('nfc.w0', 'nfc.b0', 'bf.wpc', 'bf.b0c')}
        for off, rows, cols, cp in self.fused.values():
            lab[off:off + rows * cp] = 0
        assert float(lab.max()) < n and np.all(lab == np.round(lab))
        self.index = torch.from_numpy(lab.astype(np.int64)).to(self.device)
        self.A = agg_prefix + 'agg_impl.'
        # adjoint of the gather (every parameter element sits at exactly one raw blob position): grad_flat = dblob[inverse]
        inv = np.full(n - 1, -1, np.int64)
        pos = np.nonzero(lab > 0)[0]
        inv[lab[pos].astype(np.int64) - 1] = pos
        assert (inv >= 0).all(), 'a parameter element is missing from the blob'
        self.inverse = torch.from_numpy(inv).to(self.device)
        self.shapes = [tuple(sd[k].shape) for k in self.keys]
        self.sizes = [int(np.prod(s)) if len(s) else 1 for s in self.shapes]

    def pack(self, sd):
        """sd: {key: device tensor}; returns the fp32 blob on the device (no synchronisation)."""
        flat = torch.cat([sd[k].detach().reshape(-1).to(torch.float32) for k in self.keys])
        blob = torch.cat([flat.new_zeros(1), flat])[self.index]
        d = lambda k: sd[k].detach().to(torch.float64)
        w_pe2, b_pe2 = d(self.agg_prefix + 'prob_embed.2.weight'), d(self.agg_prefix + 'prob_embed.2.bias')
        w_nf0, b_nf0 = d(self.A + 'neuray_fc.0.weight'), d(self.A + 'neuray_fc.0.bias')
        w_bfp, b_bf0 = d(self.A + 'base_fc.0.weight')[:, 175:207], d(self.A + 'base_fc.0.bias')
        vals = {'nfc.w0': (w_nf0 @ w_pe2).t(), 'nfc.b0': b_pe2 @ w_nf0.t() + b_nf0,
                'bf.wpc': (w_bfp @ w_pe2).t(), 'bf.b0c': b_pe2 @ w_bfp.t() + b_bf0}
        for name, (off, rows, cols, cp) in self.fused.items():
            assert cols == cp
            blob[off:off + rows * cp] = vals[name].reshape(-1).to(torch.float32)
        return blob

    def unpack_grad(self, dblob):
        """Adjoint of `pack` over the raw entries (= `unpack_blob_grad`, as ONE gather): blob-layout gradient -> {key: grad}.
        The fused composites are never written by the reverse kernels (their gradient arrives through the raw entries)."""
        flat = dblob.to(torch.float32)[self.inverse]
        return {k: g.reshape(s) for k, g, s in zip(self.keys, flat.split(self.sizes), self.shapes)}


def unpack_blob_grad(dblob, agg_prefix='agg_net.', dd_prefix='dist_decoder.'):
    """Adjoint of `pack_blob` over the raw (unfused) entries: a gradient in blob layout (torch tensor, any device)
    -> {reference state_dict key: gradient tensor of that parameter's shape}.  The packing is a linear map (transposes,
    the [rgb|img] <-> [img|rgb] permutation, the split of base_fc.0 into three blocks), so this is its transpose;
    tests/test_cabi.py checks <pack(w), g> == <w, unpack(g)>."""
    table = {name: (off, rows, cols, cp) for name, off, rows, cols, cp in _lib.weight_table()}
    A = agg_prefix + 'agg_impl.'
    perm = torch.as_tensor(PERM35, device=dblob.device)

    def ent(name):
        off, rows, cols, cp = table[name]
        return dblob[off:off + rows * cp].reshape(rows, cp)[:, :cols]
    out = {}

    def lin(dst_w, dst_b, key, vector=False):
        g = ent(dst_w)
        out[key + '.weight'] = g.clone() if vector else g.t().clone()
        if dst_b is not None:
            out[key + '.bias'] = ent(dst_b)[0].clone()
    lin('rd.w0', 'rd.b0', A + 'ray_dir_fc.0')
    g = ent('rd.w1')                                             # [16,35] columns in record order
    w = g.new_zeros(35, 16); w[perm] = g.t(); out[A + 'ray_dir_fc.2.weight'] = w
    b = g.new_zeros(35); b[perm] = ent('rd.b1')[0]; out[A + 'ray_dir_fc.2.bias'] = b
    for short, name in (('mean', 'mean_decoder'), ('var', 'var_decoder'), ('aw', 'aw_decoder')):
        lin(f'dd.{short}.w0', f'dd.{short}.b0', f'{dd_prefix}{name}.0')
        lin(f'dd.{short}.w2', f'dd.{short}.b2', f'{dd_prefix}{name}.2')
        lin(f'dd.{short}.w4', f'dd.{short}.b4', f'{dd_prefix}{name}.4')
    lin('pe.w0', 'pe.b0', agg_prefix + 'prob_embed.0')
    lin('pe.w2', 'pe.b2', agg_prefix + 'prob_embed.2')
    lin('nf.w0', 'nf.b0', A + 'neuray_fc.0')
    lin('nf.w2', 'nf.b2', A + 'neuray_fc.2', vector=True)
    w = dblob.new_zeros(64, 207)
    wg, wf = ent('bf.wg'), ent('bf.wf')
    for q in range(4):
        w[:, q * 35 + perm] = wg[q * 36:q * 36 + 35].t()
    w[:, 140 + perm] = wf[:35].t()
    w[:, 175:207] = ent('bf.wp').t()
    out[A + 'base_fc.0.weight'] = w
    out[A + 'base_fc.0.bias'] = ent('bf.b0')[0].clone()
    lin('bf.w2', 'bf.b2', A + 'base_fc.2')
    lin('vf.w0', 'vf.b0', A + 'vis_fc.0')
    lin('vf.w2', 'vf.b2', A + 'vis_fc.2')
    lin('v2.w0', 'v2.b0', A + 'vis_fc2.0')
    lin('v2.w2', 'v2.b2', A + 'vis_fc2.2', vector=True)
    lin('rf.w0', 'rf.b0', A + 'rgb_fc.0')
    lin('rf.w2', 'rf.b2', A + 'rgb_fc.2')
    lin('rf.w4', 'rf.b4', A + 'rgb_fc.4', vector=True)
    lin('gf.w0', 'gf.b0', A + 'geometry_fc.0')
    lin('gf.w2', 'gf.b2', A + 'geometry_fc.2')
    for dst, src in (('at.wq', 'w_qs'), ('at.wk', 'w_ks'), ('at.wv', 'w_vs'), ('at.fc', 'fc')):
        out[f'{A}ray_attention.{src}.weight'] = ent(dst).t().clone()
    out[A + 'ray_attention.layer_norm.weight'] = ent('at.ln_w')[0].clone()
    out[A + 'ray_attention.layer_norm.bias'] = ent('at.ln_b')[0].clone()
    lin('og.w0', 'og.b0', A + 'out_geometry_fc.0')
    lin('og.w1', 'og.b1', A + 'out_geometry_fc.1', vector=True)
    return out


def positional_table(n_samples, d_hid=16):
    """Sinusoid table of IBRNetWithNeuRayNeus.posenc (ibrnet.py:437-445), [n_samples, 16] fp32; the reference builds it
    in float64 numpy and casts, so it is host-side constant data rather than kernel arithmetic."""
    pos = np.arange(n_samples, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000.0, 2 * (j // 2) / d_hid)
    tab = np.where(j % 2 == 0, np.sin(ang), np.cos(ang))
    return tab.astype(np.float32)


def voxel_axis_table(resolution, volume_size=0.3):
    """utils/field_utils.py:12-25: (i * VOXEL + HALF) in python double, cast to fp32."""
    voxel = volume_size / resolution
    half = voxel / 2
    return np.array([i * voxel + half for i in range(resolution)]).astype(np.float32)


_SEED0 = {}


def seed0_model(cfg=None, seed=0):
    """The mirror `GraspNeRF` constructed under torch.manual_seed(seed): same construction order and initialisers as the
    reference (renderer.py:298-303, ibrnet.py:427-432), so seed 0 reproduces the reference's random init bit for bit
    (tests/test_boundary.py checks it against checksums of the real reference).  The caller's RNG state is preserved."""
    import torch
    from .network import name2network, NRVGN_SDF_CFG
    cfg = dict(NRVGN_SDF_CFG if cfg is None else cfg)
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        return name2network[cfg['network']](cfg).eval()


def seed0_weights(seed=0):
    """Hot-path state_dict ('agg_net.*', 'dist_decoder.*', 'fine_*' ... without the 'nr_net.' prefix) of seed0_model():
    random-init weights of the reference architecture for bench.py / smoke() / tools (no file, no dependency on tests/)."""
    if seed not in _SEED0:
        sd = seed0_model(seed=seed).state_dict()
        _SEED0[seed] = {k[len('nr_net.'):]: v.detach().clone() for k, v in sd.items()
                        if k.startswith('nr_net.') and ('agg_net.' in k or 'dist_decoder.' in k)}
    return _SEED0[seed]


# ------------------------------------------------------------------------------------------------ VGN 3-D ConvNet (K5)
_VGN_LAYERS = ('encoder.conv1', 'encoder.conv2', 'encoder.conv3', 'decoder.conv1', 'decoder.conv2', 'decoder.conv3', 'heads')


def _fold_upsampled(w):
    """Folds `nearest x2 upsample -> Conv3d(K, padding=K//2)` into 8 parity classes of 3x3x3 kernels on the low-resolution
    grid (csrc/k5_vgn_conv.cu): output voxel o = 2m + parity reads high-res taps o + d, d in [-r, r], which live in low-res
    cells m + floor((parity + d) / 2) in {m-1, m, m+1}; taps sharing a cell share their input, so their weights add up.
    w [cout, cin, K, K, K] (fp64) -> [8, cout, cin, 27], class index = px*4 + py*2 + pz."""
    import torch
    K = w.shape[-1]
    r = K // 2
    M = torch.zeros(2, 3, K, dtype=torch.float64, device=w.device)
    for par in range(2):
        for d in range(-r, r + 1):
            M[par, (par + d) // 2 + 1, d + r] = 1.0
    eff = torch.einsum('pad,qbe,rcf,oidef->pqroiabc', M, M, M, w.double())
    return eff.reshape(8, w.shape[0], w.shape[1], 27)


def pack_vgn(sd, prefix=''):
    """VGN state_dict (src/gd/networks.py:39-97 key names: encoder.conv{1,2,3}, decoder.conv{1,2,3}, conv_qual / conv_rot /
    conv_width) -> the fp32 blob gn_vgn_forward consumes.  Layout per layer comes from the library (gn_vgn_layer_info)."""
    import ctypes as C
    import torch
    from . import _lib
    lib = _lib.load()
    dev = sd[prefix + 'encoder.conv1.weight'].device
    blob = torch.zeros(lib.gn_vgn_blob_floats(), dtype=torch.float32, device=dev)
    for li, name in enumerate(_VGN_LAYERS):
        v = [C.c_int() for _ in range(7)]
        assert lib.gn_vgn_layer_info(li, *[C.byref(x) for x in v]) == 0
        cin, cout, cout_t, taps, classes, w_off, b_off = (x.value for x in v)
        if name == 'heads':
            w = torch.cat([sd[prefix + k + '.weight'] for k in ('conv_qual', 'conv_rot', 'conv_width')], 0)
            b = torch.cat([sd[prefix + k + '.bias'] for k in ('conv_qual', 'conv_rot', 'conv_width')], 0)
        else:
            w, b = sd[prefix + name + '.weight'], sd[prefix + name + '.bias']
        w, b = w.detach().double(), b.detach().double()
        assert w.shape[0] == cout and w.shape[1] == cin
        w = _fold_upsampled(w) if classes == 8 else w.reshape(1, cout, cin, taps)
        assert w.shape[-1] == taps
        nb = (cout + cout_t - 1) // cout_t
        wk = w.permute(0, 2, 3, 1).reshape(classes, cin * taps, cout)                    # [class][cin*tap][cout]
        wk = torch.nn.functional.pad(wk, (0, nb * cout_t - cout)).reshape(classes, cin * taps, nb, cout_t).permute(0, 2, 1, 3)
        blob[w_off:w_off + wk.numel()] = wk.reshape(-1).float()
        blob[b_off:b_off + cout] = b.float()
    return blob
