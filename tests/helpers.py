"""Shared helpers for the test-suite (golden loading, tolerance recipe)."""
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def golden_weights():
    import torch
    return {k: torch.from_numpy(v) for k, v in load_golden('weights_seed0.npz').items()}


def assert_close(actual, expected, rtol=1e-4, atol_scale=1e-4, what=''):
    """Tolerance recipe of SURVEY.md section 8c:  |a-b| <= rtol*|b| + atol_scale*max|b|."""
    a = np.asarray(actual, dtype=np.float64)
    b = np.asarray(expected, dtype=np.float64)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    assert np.isfinite(a).all(), f'{what}: non-finite values'
    scale = np.abs(b).max() if b.size else 0.0
    err = np.abs(a - b)
    tol = rtol * np.abs(b) + atol_scale * scale
    bad = err > tol
    rel_l2 = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
    assert not bad.any(), (f'{what}: {bad.sum()}/{bad.size} outside tolerance; max abs err {err.max():.3e} '
                           f'(scale {scale:.3e}), rel-L2 {rel_l2:.3e}')
    return rel_l2


def oracle_volume_grads(sd, scene_t, G):
    """d sum(volume*G) / d {img_feats, ray_feats, every weight} by torch autograd through the oracle (CPU).
    Returns (loss, d_img_feats, d_ray_feats, {key: grad})."""
    import torch
    from oracle import nr_oracle as O
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    sc = dict(scene_t)
    sc['img_feats'] = sc['img_feats'].detach().clone().requires_grad_(True)
    sc['ray_feats'] = sc['ray_feats'].detach().clone().requires_grad_(True)
    with torch.enable_grad():
        vol = O.sample_volume(sd, sc)
        loss = (vol * torch.as_tensor(G)).sum()
        loss.backward()
    return float(loss), sc['img_feats'].grad, sc['ray_feats'].grad, {k: v.grad for k, v in sd.items() if v.grad is not None}


RENDER_GRAD_CASE = dict(scene=dict(seed=3, num_views=4, h=96, w=160, radius=0.45), num_rays=32, qseed=7)
HOT_PREFIXES = ('agg_net.', 'fine_agg_net.', 'dist_decoder.', 'fine_dist_decoder.')


def render_grad_loss(out, g):
    """The functional behind tests/golden/render_grad_small_v4.npz (make_golden.make_render_grad_golden)."""
    import torch
    G1, G2 = torch.as_tensor(g['G1']), torch.as_tensor(g['G2'])
    pc, pf = out['pixel_colors_nr'], out['pixel_colors_nr_fine']
    G1, G2 = G1.to(pc.device).reshape(pc.shape), G2.to(pc.device).reshape(pf.shape)
    return (pc * G1).sum() + (pf * G2).sum() + 0.1 * (out['sdf_gradient_error'].sum() + out['sdf_gradient_error_fine'].sum())


def oracle_render_grads(g, dtype=None):
    """Training-mode RGB head through the oracle (torch autograd on CPU, second order for the eikonal terms):
    returns (out dict, loss, d_img_feats, d_ray_feats, {key: grad}) for the RENDER_GRAD_CASE scene."""
    import torch
    from oracle import nr_oracle as O
    from graspnerf_b200.synth import make_scene, make_query
    case = RENDER_GRAD_CASE
    scn = make_scene(**case['scene'])
    cast = (lambda t: t.to(dtype)) if dtype is not None else (lambda t: t)
    sc = {k: (cast(torch.from_numpy(v)) if isinstance(v, np.ndarray) else v) for k, v in scn.items()}
    q = make_query(scn, case['num_rays'], case['qseed'])
    que = {'coords': cast(torch.from_numpy(q['coords'][0])), 'pose': cast(torch.from_numpy(q['poses'][0])),
           'K': cast(torch.from_numpy(q['Ks'][0])), 'depth_range': cast(torch.from_numpy(q['depth_range'][0]))}
    sd = {k: cast(v).detach().clone().requires_grad_(True) for k, v in golden_weights().items()}
    sc['img_feats'] = sc['img_feats'].detach().clone().requires_grad_(True)
    sc['ray_feats'] = sc['ray_feats'].detach().clone().requires_grad_(True)
    with torch.enable_grad():
        out = O.render_rays(sd, sc, que, 40, 40, u=cast(torch.from_numpy(g['u'][0])), train=True,
                            fine_depth=cast(torch.from_numpy(g['depth_fine'][0])))
        loss = render_grad_loss({k: v for k, v in out.items()}, {'G1': cast(torch.as_tensor(g['G1'])), 'G2': cast(torch.as_tensor(g['G2']))})
        loss.backward()
    return out, float(loss), sc['img_feats'].grad, sc['ray_feats'].grad, {k: v.grad for k, v in sd.items() if v.grad is not None}


def grad_close(a, b, tol_max, tol_l2, what, abs_floor=2e-6):
    """|a-b|_max <= tol_max * max|b| and rel-L2 <= tol_l2 (gradient tensors; atomics / fp32 order make element-wise
    relative checks meaningless near zero).  Gradients that are analytically zero (e.g. the bias in front of a softmax)
    are fp32 noise on both sides: they pass when |a-b|_max <= abs_floor."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    scale = max(np.abs(b).max(), 1e-30)
    if np.isfinite(a).all() and np.abs(a - b).max() <= abs_floor:
        return 0.0, 0.0
    emax = np.abs(a - b).max() / scale
    el2 = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
    assert np.isfinite(a).all() and emax <= tol_max and el2 <= tol_l2, f'{what}: max err {emax:.3e} (tol {tol_max}), rel-L2 {el2:.3e} (tol {tol_l2})'
    return emax, el2
