"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/graspnerf_b200.h declares;
host-side packing logic is consistent with the library's own weight table.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.helpers import ROOT, golden_weights


@pytest.fixture(scope='module')
def lib():
    from graspnerf_b200.build import build_library
    from graspnerf_b200 import _lib
    build_library()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'graspnerf_b200.h')).read()
    names = set(re.findall(r'\b(gn_[a-z0-9_]+)\s*\(', hdr))
    assert len(names) >= 12
    for n in sorted(names):
        assert hasattr(lib, n), f'{n} declared in the header but not exported'
    assert b'sm_100a' in lib.gn_version()


def test_struct_sizes_match(lib):
    from graspnerf_b200 import _lib
    assert lib.gn_sizeof_k1_params() == ctypes.sizeof(_lib.GnK1Params)
    assert lib.gn_sizeof_k2a_params() == ctypes.sizeof(_lib.GnK2aParams)
    assert lib.gn_sizeof_k2b_params() == ctypes.sizeof(_lib.GnK2bParams)
    assert lib.gn_sizeof_k3_params() == ctypes.sizeof(_lib.GnK3Params)


def test_weight_table_is_dense_and_aligned(lib):
    from graspnerf_b200 import _lib
    tab = _lib.weight_table()
    off = 0
    for name, o, rows, cols, cp in tab:
        assert o == off and o % 4 == 0 and cp % 4 == 0 and cp >= cols, name
        off += rows * cp
    assert off == lib.gn_weight_blob_floats()


def test_pack_blob_roundtrip(lib):
    """Every reference tensor lands where the table says, k-major, with the [rgb|img] -> [img|rgb] permutation."""
    from graspnerf_b200 import _lib
    from graspnerf_b200.weights import pack_blob, PERM35
    sd = golden_weights()
    blob = pack_blob(sd, 'fine_agg_net.', 'fine_dist_decoder.')
    tab = {n: (o, r, c, cp) for n, o, r, c, cp in _lib.weight_table()}

    def entry(n):
        o, r, c, cp = tab[n]
        return blob[o:o + r * cp].reshape(r, cp)[:, :c]
    A = 'fine_agg_net.agg_impl.'
    w = sd[A + 'base_fc.0.weight'].numpy()
    assert np.array_equal(entry('bf.wp'), w[:, 175:].T)
    wg = entry('bf.wg')
    assert np.array_equal(wg[36 + 0], w[:, 35 + 3]) and np.array_equal(wg[36 + 32], w[:, 35 + 0]) and not wg[35].any()
    assert np.array_equal(entry('dd.var.w4'), sd['fine_dist_decoder.var_decoder.4.weight'].numpy().T)
    assert np.array_equal(entry('at.fc'), sd[A + 'ray_attention.fc.weight'].numpy().T)
    assert np.array_equal(entry('og.w1'), sd[A + 'out_geometry_fc.1.weight'].numpy())
    assert np.array_equal(entry('rd.w1'), sd[A + 'ray_dir_fc.2.weight'].numpy()[PERM35].T)
    assert np.array_equal(entry('rd.w0'), sd[A + 'ray_dir_fc.0.weight'].numpy().T)


def test_product_path_refuses_cpu():
    """No CPU fallback: the torch front end raises on host tensors."""
    from graspnerf_b200 import ops
    t = torch.zeros(2, 3, 8, 8)
    with pytest.raises(RuntimeError):
        ops.Scene(t, torch.zeros(2, 32, 2, 2), torch.zeros(2, 32, 2, 2), torch.zeros(2, 3, 4), torch.zeros(2, 3, 3), torch.zeros(2, 2))


def test_unpack_blob_grad_is_the_adjoint_of_pack_blob(lib):
    """<pack(w), g> == <w, unpack(g)> for random w, g: the weight gradients the backward kernels produce in blob layout map
    back onto the reference's parameters by the exact transpose of the packing (permutations, transposes, block split)."""
    import torch
    from graspnerf_b200 import _lib
    from graspnerf_b200.weights import pack_blob, unpack_blob_grad
    from tests.helpers import golden_weights
    rng = np.random.default_rng(0)
    w = {k: torch.from_numpy(rng.standard_normal(v.shape).astype(np.float32)) for k, v in golden_weights().items()
         if k.startswith(('agg_net.', 'dist_decoder.'))}
    blob = pack_blob(w)
    g = rng.standard_normal(blob.shape).astype(np.float32)
    for name, off, rows, cols, cp in _lib.weight_table():          # entries the backward kernels never write
        if name.startswith('nfc.') or name in ('bf.wpc', 'bf.b0c'):
            g[off:off + rows * cp] = 0
    ug = unpack_blob_grad(torch.from_numpy(g))
    lhs = float(np.dot(blob.astype(np.float64), g.astype(np.float64)))
    rhs = sum(float((w[k].double() * ug[k].double()).sum()) for k in ug)
    assert abs(lhs - rhs) <= 1e-9 * max(1.0, abs(lhs))
    assert all(ug[k].shape == w[k].shape for k in ug) and len(ug) == 62


def test_device_packer_equals_host_packer(lib):
    """weights.DevicePacker (gather index + fp64 torch matmuls for the fused composites; what training uses every step)
    produces the blob of weights.pack_blob: raw entries bit for bit, composites to one fp32 ulp."""
    import torch
    from graspnerf_b200.weights import pack_blob, DevicePacker
    from tests.helpers import golden_weights
    sd = golden_weights()
    for agg, dd in (('agg_net.', 'dist_decoder.'), ('fine_agg_net.', 'fine_dist_decoder.')):
        ref = pack_blob(sd, agg, dd)
        got = DevicePacker(sd, agg, dd, 'cpu').pack(sd).numpy()
        assert got.shape == ref.shape
        fused = np.zeros(ref.shape, bool)
        for name, off, rows, cols, cp in __import__('graspnerf_b200')._lib.weight_table():
            if name in ('nfc.w0', 'nfc.b0', 'bf.wpc', 'bf.b0c'):
                fused[off:off + rows * cp] = True
        assert np.array_equal(got[~fused], ref[~fused])
        assert np.allclose(got[fused], ref[fused], rtol=2e-7, atol=1e-9) and np.abs(ref[fused]).sum() > 0
        # the one-gather adjoint equals weights.unpack_blob_grad (the transposed packing) entry by entry
        from graspnerf_b200.weights import unpack_blob_grad
        g = torch.from_numpy(np.random.default_rng(1).standard_normal(ref.shape).astype(np.float32))
        a, b = DevicePacker(sd, agg, dd, 'cpu').unpack_grad(g), unpack_blob_grad(g, agg, dd)
        assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_launchers_reject_bad_arguments_before_touching_cuda(lib):
    """Error convention of the C ABI (include/graspnerf_b200.h): 0 ok, < 0 argument error (checked before any CUDA call, so
    this runs without a GPU), > 0 cudaError_t.  Nothing is launched here."""
    from graspnerf_b200 import _lib
    p = _lib.GnK1Params(); p.B, p.V, p.N = 1, 0, 10
    assert lib.gn_k1_forward(ctypes.byref(p), None) == -1                       # V out of range
    p.V, p.volume_mode, p.R, p.N = 6, 1, 40, 123
    assert lib.gn_k1_forward(ctypes.byref(p), None) == -3                       # N != R^3 / missing tables
    p.volume_mode, p.N = 0, 64
    assert lib.gn_k1_forward(ctypes.byref(p), None) == -4                       # ray mode without pts / que_dir
    a = _lib.GnK2aParams(); a.B, a.N, a.V = 1, 10, 40
    assert lib.gn_k2a_forward(ctypes.byref(a), None) == -1 and lib.gn_k2a_forward_tc(ctypes.byref(a), None) < 0
    b = _lib.GnK2bBwdParams(); b.B, b.N, b.dn = 1, 64000, 0
    assert lib.gn_k2b_backward(ctypes.byref(b), None) == -1
    b.dn = 40
    assert lib.gn_k2b_backward(ctypes.byref(b), None) == -2                     # NULL buffers
    c = _lib.GnK2aBwdParams(); c.B, c.N, c.V = 1, 10, 6
    assert lib.gn_k2a_backward(ctypes.byref(c), None) == -2
    d = _lib.GnK1BwdParams(); d.B, d.N, d.V = 1, 10, 6
    assert lib.gn_k1_backward(ctypes.byref(d), None) == -2
    assert lib.gn_sizeof_k2b_bwd_params() == ctypes.sizeof(_lib.GnK2bBwdParams)
    assert lib.gn_sizeof_k2a_bwd_params() == ctypes.sizeof(_lib.GnK2aBwdParams)
    assert lib.gn_sizeof_k1_bwd_params() == ctypes.sizeof(_lib.GnK1BwdParams)


def test_sass_contains_the_blackwell_instructions_the_design_claims():
    """Static check on the built objects (cuobjdump, no GPU needed): the tensor-core K2a kernels really issue tcgen05 MMAs from
    tensor memory (UTCHMMA, LDTM / STTM), commit on mbarriers (UTCBAR) and fetch their weights with bulk async copies (UBLKCP);
    """
    import shutil
    import subprocess
    from graspnerf_b200.build import LIBDIR
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')

    def sass(obj):
        path = os.path.join(LIBDIR, obj)
        if not os.path.exists(path):
            pytest.skip(f'{obj} not built in-tree')
        return subprocess.run([cuobjdump, '-sass', path], capture_output=True, text=True).stdout
    for obj in ('k2a_head_tc3.o',):
        s = sass(obj)
        assert s.count('UTCHMMA') >= 100 and 'LDTM' in s and 'STTM' in s and 'UTCBAR' in s and 'UBLKCP.S.G' in s, obj
        assert 'sm_100a' in s
