"""graspnerf_b200.train restates the four losses of nrvgn_sdf.yaml (render, depth, sdf + eikonal, vgn) for environments
without the reference's loss dependencies.  tests/golden/losses.json holds the values of the reference's OWN loss classes
(network/loss.py) on a seeded random case (make_golden.py --losses); the restatements must reproduce them."""
import json
import os

import numpy as np
import torch

from tests.helpers import GOLDEN


def _case(seed=5, rn=24, pn=64, rfn=3, h=20, w=28, G=16, R=8):
    """Same generator as tests/golden/make_golden.py:loss_case (numpy PCG64, same call order)."""
    rng = np.random.default_rng(seed)
    f = lambda *s: torch.from_numpy(rng.random(s, dtype=np.float32))
    quat = rng.standard_normal((G, 2, 4)).astype(np.float32); quat /= np.linalg.norm(quat, axis=-1, keepdims=True)
    qp = rng.standard_normal((G, 4)).astype(np.float32); qp /= np.linalg.norm(qp, axis=-1, keepdims=True)
    data_pr = {'pixel_colors_gt': f(1, rn, 3), 'pixel_colors_nr': f(1, rn, 3), 'pixel_colors_nr_fine': f(1, rn, 3),
               'ray_mask': torch.from_numpy(rng.random((1, rn)) < 0.7),
               'depth_coords': torch.from_numpy(np.stack([rng.integers(0, h, (rfn, pn)), rng.integers(0, w, (rfn, pn))], -1)),
               'depth_mean': f(rfn, pn), 'depth_mean_fine': f(rfn, pn),
               'volume': f(1, 1, R, R, R) * 2 - 1, 'sdf_gradient_error': f(1, 2), 's': f(1, 1),
               'vgn_pred': (f(G) * 0.98 + 0.01, torch.from_numpy(qp), f(G) * 10)}
    sdf_gt = f(R, R, R) * 2 - 1
    sdf_gt[0, 0, :3] = -1.0
    data_gt = {'scene_name': 'vgn_syn/0', 'ref_imgs_info': {'true_depth': f(rfn, 1, h, w) * 0.6 + 0.2,
               'depth_range': torch.tensor([[0.2, 0.8]] * rfn), 'sdf_gt': sdf_gt},
               'grasp_info': [torch.from_numpy(rng.integers(0, R, (G, 3))), torch.from_numpy((rng.random(G) < 0.5).astype(np.float32)),
                              torch.from_numpy(quat), f(G) * 10]}
    return data_pr, data_gt


def test_restated_losses_match_the_reference_loss_classes():
    from graspnerf_b200 import train
    g = json.load(open(os.path.join(GOLDEN, 'losses.json')))
    pr, gt = _case()
    close = lambda a, b: abs(float(a) - b) <= 1e-6 + 2e-6 * abs(b)
    assert close(train.render_loss(pr, fine=False), g['loss_rgb_nr'])
    assert close(train.render_loss(pr, fine=True), g['loss_rgb_nr'] + g['loss_rgb_nr_fine'])
    assert close(train.eikonal_loss(pr), g['loss_eikonal'])
    assert close(train.sdf_loss(pr['volume'], gt['ref_imgs_info']['sdf_gt']), g['loss_sdf'])
    assert close(train.vgn_loss(pr['vgn_pred'], gt['grasp_info']), g['loss_vgn'])
    assert close(train.depth_loss(pr, gt['ref_imgs_info']), g['loss_depth'] + g['loss_depth_fine'])
    total = train.training_losses(pr, gt)
    want = sum(g[k] for k in ('loss_rgb_nr', 'loss_rgb_nr_fine', 'loss_depth', 'loss_depth_fine', 'loss_sdf', 'loss_eikonal', 'loss_vgn'))
    assert close(total, want), (float(total), want)       # Trainer sums every key that starts with 'loss' (trainer.py:152-155)
