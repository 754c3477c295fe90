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
