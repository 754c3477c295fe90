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
