"""The torch-side modules of the mirror (2-D encoders, depth-mean head, VGN 3-D conv: SURVEY.md section 8f, kept in
PyTorch/cuDNN by design) against the UNMODIFIED reference, live, on CPU.  Runs only where /root/reference exists (the authoring
container, where the driver runs the CPU suite); skipped on the GPU box.  The hot path itself is covered by the fixtures."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from ref_harness import reference_available, build_reference_net      # noqa: E402
from graspnerf_b200.synth import make_scene                            # noqa: E402
from tests.test_boundary import CFG                                    # noqa: E402

pytestmark = pytest.mark.skipif(not reference_available(), reason='needs the reference checkout (/root/reference)')


def test_torch_side_modules_match_the_live_reference():
    from graspnerf_b200.network import name2network
    orig_cuda, orig_to = torch.Tensor.cuda, torch.Tensor.to   # the harness shims these two for the reference's hard-coded .cuda()
    try:
        _, ref_net = build_reference_net(0)                   # torch.manual_seed(0) inside
    finally:
        torch.Tensor.cuda, torch.Tensor.to = orig_cuda, orig_to
    torch.manual_seed(0)
    net = name2network[CFG['network']](dict(CFG)).eval()
    sd_ref = ref_net.state_dict()
    assert all(torch.equal(v, sd_ref[k]) for k, v in net.state_dict().items())       # same init -> same weights (also test_boundary)
    scene = make_scene(seed=3, num_views=4, h=96, w=160, radius=0.45)
    imgs = torch.from_numpy(scene['imgs'])
    ref = {'imgs': imgs}
    with torch.no_grad():
        # image_encoder / init_net / vis_encoder (renderer.py:275-279)
        a_img, b_img = net.nr_net.image_encoder(imgs), ref_net.nr_net.image_encoder(imgs)
        a_ray = net.nr_net.vis_encoder(net.nr_net.init_net(ref, ref, False), a_img)
        b_ray = ref_net.nr_net.vis_encoder(ref_net.nr_net.init_net(ref, ref, False), b_img)
        assert a_img.shape == (4, 32, 24, 40) and torch.allclose(a_img, b_img, rtol=1e-5, atol=1e-6)
        assert torch.allclose(a_ray, b_ray, rtol=1e-5, atol=1e-6)
        # depth-mean head (renderer.py:222-266): same pixels under the same RNG state, same decoders
        info = {'imgs': imgs, 'ray_feats': b_ray}
        torch.manual_seed(7)
        da = net.nr_net.predict_mean_for_depth_loss(info)
        torch.manual_seed(7)
        db = ref_net.nr_net.predict_mean_for_depth_loss(info)
        assert set(da) == set(db)
        assert torch.equal(da['depth_coords'], db['depth_coords'])
        for k in ('depth_mean', 'depth_mean_2', 'depth_mean_fine', 'depth_mean_fine_2'):
            assert torch.allclose(da[k], db[k], rtol=1e-5, atol=1e-6), k
        # VGN head on a volume (gd/networks.py:39-97, renderer.py:323-330)
        vol = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (1, 1, 40, 40, 40)).astype(np.float32))
        for x, y in zip(net.vgn_net(vol), ref_net.vgn_net(vol)):
            assert torch.allclose(x, y, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('kw', [dict(seed=41, num_views=5, h=80, w=112, radius=0.4), dict(seed=42, num_views=2, h=64, w=64, radius=0.25, theta=1.0)])
def test_oracle_sample_volume_matches_the_live_reference_on_fresh_scenes(kw):
    """Beyond the committed fixtures: the oracle against the reference's sample_volume on scenes that are in no fixture
    (other view counts / image sizes / camera rings, incl. a 2-view close-up with many invalid projections)."""
    from oracle import nr_oracle as O
    from tests.helpers import assert_close
    orig_cuda, orig_to = torch.Tensor.cuda, torch.Tensor.to
    try:
        _, ref_net = build_reference_net(0)
        scene = make_scene(**kw)
        sc = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in scene.items()}
        with torch.no_grad():
            want = ref_net.nr_net.sample_volume(dict(sc))
    finally:
        torch.Tensor.cuda, torch.Tensor.to = orig_cuda, orig_to
    sd = {k: v for k, v in ref_net.nr_net.state_dict().items()}
    got = O.sample_volume(sd, sc)
    rel = assert_close(got, want, what=f'volume {kw}')
    assert rel < 1e-5


def test_oracle_render_rays_matches_the_live_reference_on_a_fresh_scene():
    """RGB head (coarse + fine, eval) on a scene / query that is in no fixture.  The fine pass gets the reference's own fine
    sample depths (the inverse-CDF sampler is ill-conditioned where the coarse hit probability vanishes, see make_golden.py)."""
    from oracle import nr_oracle as O
    from graspnerf_b200.synth import make_query
    from tests.helpers import assert_close
    kw = dict(seed=43, num_views=3, h=64, w=96, radius=0.5)
    scene = make_scene(**kw)
    sc = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in scene.items()}
    q = make_query(scene, 20, 5)
    que_ref = {k: torch.from_numpy(v) for k, v in q.items()}
    orig_cuda, orig_to = torch.Tensor.cuda, torch.Tensor.to
    try:
        _, ref_net = build_reference_net(0)
        nr = ref_net.nr_net
        res = nr.render_impl(que_ref, dict(sc), False)
        from network.render_ops import sample_depth, sample_fine_depth
        with torch.no_grad():
            depth, _ = sample_depth(que_ref['depth_range'], que_ref['coords'], nr.cfg['depth_sample_num'], False)
            fd = sample_fine_depth(depth, res['hit_prob_nr'].detach(), que_ref['depth_range'], nr.cfg['fine_depth_sample_num'], False)
    finally:
        torch.Tensor.cuda, torch.Tensor.to = orig_cuda, orig_to
    sd = dict(nr.state_dict())
    que = {'coords': que_ref['coords'][0], 'pose': que_ref['poses'][0], 'K': que_ref['Ks'][0], 'depth_range': que_ref['depth_range'][0]}
    out = O.render_rays(sd, sc, que, 40, 40, fine_depth=torch.sort(fd, -1)[0][0])
    for k in ('sdf_values', 'alpha_values', 'hit_prob_nr', 'colors_nr', 'pixel_colors_nr', 'render_depth'):
        for sfx in ('', '_fine'):
            assert_close(out[k + sfx].detach(), res[k + sfx][0].detach(), what=k + sfx)
    assert torch.equal(out['ray_mask'], res['ray_mask'][0]) and torch.equal(out['ray_mask_fine'], res['ray_mask_fine'][0])


def test_install_swaps_the_reference_registry():
    """graspnerf_b200.install(): the reference's own registry (renderer.py:333-335) hands out the mirror class, which loads a
    reference state_dict unchanged - what train.sh / sim_grasp.py need to run with zero source edits."""
    import graspnerf_b200
    from graspnerf_b200.network import GraspNeRF as Mirror
    orig_cuda, orig_to = torch.Tensor.cuda, torch.Tensor.to
    try:
        cfg, ref_net = build_reference_net(0)
        import network.renderer as ref_mod
        saved = (ref_mod.name2network['grasp_nerf'], ref_mod.GraspNeRF, ref_mod.NeuralRayRenderer)
        try:
            reg = graspnerf_b200.install()
            assert reg is ref_mod.name2network and reg['grasp_nerf'] is Mirror
            net = reg[cfg['network']](cfg)                     # the reference's cfg (yaml) builds the mirror
            missing, unexpected = net.load_state_dict(ref_net.state_dict(), strict=True)
            assert not missing and not unexpected
        finally:
            ref_mod.name2network['grasp_nerf'], ref_mod.GraspNeRF, ref_mod.NeuralRayRenderer = saved
            ref_mod.__graspnerf_b200__ = False
    finally:
        torch.Tensor.cuda, torch.Tensor.to = orig_cuda, orig_to
