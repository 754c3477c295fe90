"""The host-side mirror of the reference's model API: state_dict layout (CPU) and a full GraspNeRF.forward on the GPU
against the reference's own output (tests/golden/forward_small_v4.npz, made by tests/golden/make_golden.py --model)."""
import json
import os

import numpy as np
import pytest
import torch
import yaml

from tests.helpers import GOLDEN, load_golden, assert_close
from graspnerf_b200.synth import make_scene, make_query

# the shipped config (src/nr/configs/nrvgn_sdf.yaml:5-31), network-relevant keys only
CFG = yaml.safe_load("""
network: grasp_nerf
init_net_type: cost_volume
agg_net_type: neus
use_hierarchical_sampling: true
use_depth_loss: true
dist_decoder_cfg: {use_vis: false}
fine_dist_decoder_cfg: {use_vis: false}
ray_batch_num: 4096
sample_volume: true
render_rgb: true
volume_type: [sdf]
volume_resolution: 40
depth_sample_num: 40
fine_depth_sample_num: 40
agg_net_cfg: {sample_num: 40, init_s: 0.3, fix_s: 0}
fine_agg_net_cfg: {sample_num: 40, init_s: 0.3, fix_s: 0}
render_depth: true
""")


def _build():
    from graspnerf_b200.network import name2network
    torch.manual_seed(0)
    return name2network[CFG['network']](dict(CFG)).eval()


def test_state_dict_matches_reference_layout_and_seed0_init():
    """Same 348 keys and shapes as the reference; constructed in the same order with the same initialisers, so
    torch.manual_seed(0) reproduces the reference's weights exactly (checksums from the real reference)."""
    table = json.load(open(os.path.join(GOLDEN, 'state_dict_keys.json')))
    sd = _build().state_dict()
    assert set(sd) == set(table), sorted(set(sd) ^ set(table))[:10]
    for k, (shape, s, a) in table.items():
        assert list(sd[k].shape) == shape, k
        assert abs(float(sd[k].double().sum()) - s) <= 1e-9 * max(1.0, a), f'{k}: seed-0 init differs from the reference'
    hot = load_golden('weights_seed0.npz')
    for k, v in hot.items():
        assert np.array_equal(sd['nr_net.' + k].numpy(), v), k


def test_hot_path_has_no_cpu_fallback():
    """The mirror's hot path raises on host tensors instead of falling back to torch ops (ops._require_cuda)."""
    from graspnerf_b200 import ops
    z = torch.zeros(2, 3, 8, 8)
    with pytest.raises(RuntimeError, match='no CPU path'):
        ops.Scene(z, torch.zeros(2, 32, 2, 2), torch.zeros(2, 32, 2, 2), torch.zeros(2, 3, 4), torch.zeros(2, 3, 3), torch.zeros(2, 2))


@pytest.mark.gpu
def test_full_forward_matches_reference_output():
    """GraspNeRF.forward (eval, render_rgb off like main.py:150) on the GPU vs the reference's CPU output.  The 2-D
    encoders run in cuDNN fp32 (TF32 off), so the comparison carries conv round-off: tolerance 2e-3."""
    g = load_golden('forward_small_v4.npz')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device('cuda:0')
    net = _build().to(dev)
    net.nr_net.cfg['render_rgb'] = False
    scene = make_scene(seed=3, num_views=4, h=96, w=160, radius=0.45)
    ref = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in scene.items()
           if k not in ('img_feats', 'ray_feats')}
    q = {k: torch.from_numpy(v).to(dev) for k, v in make_query(scene, 16, 7).items()}
    data = {'step': 0, 'eval': True, 'full_vol': True, 'ref_imgs_info': ref, 'que_imgs_info': q, 'src_imgs_info': ref}
    with torch.no_grad():
        out = net(data)
    for k in ('volume', 'vgn_pred', 'depth_mean', 'depth_coords', 'depth_mean_2', 'depth_mean_fine', 'depth_mean_fine_2'):
        assert k in out
    assert out['volume'].shape == (1, 1, 40, 40, 40)
    assert_close(out['volume'][0, 0].cpu(), g['volume'], rtol=2e-3, atol_scale=2e-3, what='volume vs reference forward')
    assert_close(out['vgn_pred'][0][0, 0].cpu(), g['qual'], rtol=2e-3, atol_scale=2e-3, what='vgn quality volume')


@pytest.mark.gpu
def test_render_rgb_forward_keys():
    """forward with render_rgb on emits the reference's key set (renderer.py:90-138,160-161)."""
    dev = torch.device('cuda:0')
    net = _build().to(dev)
    scene = make_scene(seed=3, num_views=4, h=96, w=160, radius=0.45)
    ref = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in scene.items()
           if k not in ('img_feats', 'ray_feats')}
    q = {k: torch.from_numpy(v).to(dev) for k, v in make_query(scene, 32, 7).items()}
    with torch.no_grad():
        out = net.nr_net({'step': 0, 'eval': True, 'ref_imgs_info': ref, 'que_imgs_info': q})
    for k in ('pixel_colors_nr', 'pixel_colors_gt', 'hit_prob_nr', 'alpha_values', 'colors_nr', 'sdf_values',
              'sdf_gradient_error', 's', 'ray_mask', 'render_depth'):
        assert k in out and (k + '_fine') in out, k
    assert out['pixel_colors_nr_fine'].shape == (1, 32, 3) and out['ray_mask'].shape == (1, 32)
    assert torch.isfinite(out['pixel_colors_nr_fine']).all()
