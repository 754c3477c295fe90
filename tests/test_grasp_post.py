"""Grasp post-processing (`process` / `select`, src/nr/main.py:23-74): the oracle against the reference's own source
(CPU, where /root/reference exists) and gn_k4_grasp_post against the oracle (GPU): filtered volume bit-exact, grasp list in
np.argwhere order."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import grasp_post as G
from oracle.ref_harness import REF_ROOT

MAIN_PY = os.path.join('/root/reference', 'src', 'nr', 'main.py')


def _volumes(seed, R=40, smooth=False):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    tsdf = rng.uniform(-1, 1, (R, R, R)).astype(np.float32)
    if smooth:                                            # a surface-like TSDF: large connected inside / outside regions
        tsdf = np.clip(ndimage.gaussian_filter(tsdf, 3.0) * 12, -1, 1).astype(np.float32)
    qual = (rng.uniform(0, 1, (R, R, R)) ** 0.15).astype(np.float32)
    rot = rng.standard_normal((4, R, R, R)).astype(np.float32)
    rot /= np.linalg.norm(rot, axis=0, keepdims=True)
    width = rng.uniform(0, 11, (R, R, R)).astype(np.float32)
    return tsdf, qual, rot, width


@pytest.mark.skipif(not os.path.exists(MAIN_PY), reason='needs the reference checkout')
def test_oracle_is_the_references_process_and_select():
    """Executes the UNMODIFIED source text of `process` and `select` (main.py:23-74; the module itself cannot be imported:
    skimage / pybullet-side imports) with numpy + scipy.ndimage and a stub for select_index, and compares with the oracle."""
    from scipy import ndimage
    src = open(MAIN_PY).read()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ('process', 'select')]
    ns = {'np': np, 'ndimage': ndimage,
          'select_index': lambda q, r, w, index: ((tuple(index), r[:, index[0], index[1], index[2]], w[tuple(index)]), q[tuple(index)])}
    exec(compile(ast.Module(body=fns, type_ignores=[]), MAIN_PY, 'exec'), ns)
    for seed, smooth in ((0, False), (1, True)):
        tsdf, qual, rot, width = _volumes(seed, smooth=smooth)
        kw = dict(tsdf_thres_high=0.0, tsdf_thres_low=-0.85)               # main.py:92-93
        q_ref, r_ref, w_ref = ns['process'](tsdf[None, None], qual.copy()[None, None], rot[None], width[None, None], **kw)
        q_or, r_or, w_or = G.process(tsdf[None, None], qual.copy()[None, None], rot[None], width[None, None], **kw)
        assert np.array_equal(q_ref, q_or)
        grasps, scores, indexs = ns['select'](q_ref.copy(), r_ref, w_ref)
        idx, sc, ro, wi = G.select(q_or, r_or, w_or)
        assert len(grasps) == len(idx) > 0
        assert np.array_equal(np.asarray(indexs), idx) and np.array_equal(np.asarray(scores, np.float32), sc.astype(np.float32))
        assert all(np.array_equal(g[1], ro[i]) and g[2] == wi[i] for i, g in enumerate(grasps))


@pytest.mark.gpu
@pytest.mark.parametrize('case', [dict(seed=0, smooth=False, thr=0.9), dict(seed=1, smooth=True, thr=0.9),
                                  dict(seed=2, smooth=True, thr=0.5, R=24, size=3, sigma=1.5)])
def test_grasp_post_kernel_matches_oracle(case):
    from graspnerf_b200 import ops
    R = case.get('R', 40)
    tsdf, qual, rot, width = _volumes(case['seed'], R=R, smooth=case['smooth'])
    kw = dict(tsdf_thres_high=0.0, tsdf_thres_low=-0.85, gaussian_filter_sigma=case.get('sigma', 1.0))
    q_or, r_or, w_or = G.process(tsdf, qual.copy(), rot, width, **kw)
    idx, sc, ro, wi = G.select(q_or, r_or, w_or, threshold=case['thr'], max_filter_size=case.get('size', 4))
    dev = torch.device('cuda:0')
    t = [torch.from_numpy(x).to(dev) for x in (tsdf, qual, rot, width)]
    q_gpu, grasps, count = ops.grasp_post(*t, threshold=case['thr'], max_filter_size=case.get('size', 4), max_grasps=1024, **kw)
    torch.cuda.synchronize()
    assert np.array_equal(q_gpu.cpu().numpy(), q_or), 'process(): filtered + masked quality volume must be bit-identical to scipy'
    n = int(count.item())
    assert n == len(idx) and n > 0
    g = grasps[:n].cpu().numpy()
    assert np.array_equal(g[:, :3].astype(np.int64), idx), 'np.argwhere order'
    assert np.array_equal(g[:, 3], sc.astype(np.float32)) and np.array_equal(g[:, 4:8], ro) and np.array_equal(g[:, 8], wi)


@pytest.mark.gpu
def test_grasp_post_truncates_at_max_grasps():
    from graspnerf_b200 import ops
    tsdf, qual, rot, width = _volumes(3, smooth=True)
    dev = torch.device('cuda:0')
    t = [torch.from_numpy(x).to(dev) for x in (tsdf, qual, rot, width)]
    _, g_all, c_all = ops.grasp_post(*t, threshold=0.5, tsdf_thres_high=0.0, tsdf_thres_low=-0.85, max_grasps=2048)
    _, g_few, c_few = ops.grasp_post(*t, threshold=0.5, tsdf_thres_high=0.0, tsdf_thres_low=-0.85, max_grasps=5)
    assert int(c_all.item()) == int(c_few.item()) > 5                      # count reports all, rows are truncated
    assert torch.equal(g_few, g_all[:5])
