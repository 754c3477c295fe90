"""The mirror's stage-by-stage API (same function / method names as the reference: render_ops.py and renderer.py:62-162,
aggregate_net.py:125) on the GPU against the UNMODIFIED reference running the same calls on the CPU.  The reference comes
from oracle/_ref (verbatim copy made by oracle/make_ref.py; it travels to the GPU box) or /root/reference; skipped when
neither exists."""
import numpy as np
import pytest
import torch

from tests.helpers import assert_close
from oracle import ref_harness as RH
from graspnerf_b200.synth import make_scene, make_query
from graspnerf_b200.weights import seed0_model

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not RH.reference_available(), reason='needs a copy of the reference (oracle/_ref)')]
DEV = 'cuda:0'


def test_stage_api_matches_the_reference_stage_by_stage():
    from graspnerf_b200.network import render_ops as M
    scn = make_scene(seed=17, num_views=4, h=96, w=160, radius=0.5)
    q = make_query(scn, 24, 5)
    ref_info = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in scn.items()}
    que = {k: torch.from_numpy(q[k]) for k in ('coords', 'poses', 'Ks', 'depth_range')}
    # ---------------- the reference, on the CPU
    with RH.shims():
        cfg, ref_net = RH.build_reference_net(0)
        import network.render_ops as R
        nr = ref_net.nr_net
        r_depth, _ = R.sample_depth(que['depth_range'], que['coords'], 40, False)
        r_inv = R.depth2inv_dists(r_depth, que['depth_range'])
        r_pts, r_dir = R.depth2points(que, r_depth)
        r_centers, r_dirs = R.coords2rays(que['coords'], que['poses'], que['Ks'])
        with torch.no_grad():
            r_prj = R.project_points_dict(ref_info, r_pts)
            r_prj = nr.predict_proj_ray_prob(r_prj, ref_info, r_inv, False)
            r_prj = nr.get_img_feats(ref_info, r_prj)
        r_out = nr.network_rendering(r_prj, r_dir, r_pts, r_depth, False, False, is_sdf=True)    # autograd.grad inside: grad mode on
        r_out = {k: v.detach() for k, v in r_out.items()}
        r_pts, r_dir = r_pts.detach(), r_dir.detach()           # ibrnet.py:486 switches requires_grad on in place
        r_prj = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in r_prj.items()}
    # ---------------- the mirror, on the GPU, same calls
    net = seed0_model().to(DEV).eval()
    mnr = net.nr_net
    g_ref = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in ref_info.items()}
    g_que = {k: v.to(DEV) for k, v in que.items()}
    with torch.no_grad():
        m_depth, _ = M.sample_depth(g_que['depth_range'], g_que['coords'], 40, False)
        m_inv = M.depth2inv_dists(m_depth, g_que['depth_range'])
        m_pts, m_dir = M.depth2points(g_que, m_depth)
        m_centers, m_dirs = M.coords2rays(g_que['coords'], g_que['poses'], g_que['Ks'])
        m_prj = M.project_points_dict(g_ref, m_pts)
        m_prj = mnr.predict_proj_ray_prob(m_prj, g_ref, m_inv, False)
        m_prj = mnr.get_img_feats(g_ref, m_prj)
        m_out = mnr.network_rendering(m_prj, m_dir, m_pts, m_depth, False, False, is_sdf=True)
    torch.cuda.synchronize()
    assert torch.equal(m_depth.cpu(), r_depth), 'coarse depth table'
    assert_close(m_centers.cpu(), r_centers, rtol=1e-5, atol_scale=1e-6, what='coords2rays centers')
    assert_close(m_dirs.cpu(), r_dirs, rtol=1e-4, atol_scale=1e-5, what='coords2rays directions')
    assert_close(m_pts.cpu(), r_pts, rtol=1e-5, atol_scale=1e-6, what='depth2points que_pts')
    assert_close(m_dir.cpu(), r_dir, rtol=1e-5, atol_scale=1e-6, what='depth2points que_dir')
    assert_close(m_inv.cpu()[..., :-1], r_inv[..., :-1], rtol=1e-4, atol_scale=1e-5, what='depth2inv_dists')
    assert torch.equal(m_prj['mask'].cpu(), r_prj['mask']), 'validity mask table'
    for k in ('dir', 'depth', 'ray_feats', 'rgb', 'img_feats', 'hit_prob', 'vis'):
        a, b = m_prj[k].cpu(), r_prj[k]
        if k in ('dir', 'depth'):                       # the reference keeps dir / depth of invalid projections; the record masks nothing there either
            pass
        assert_close(a, b, what=f'prj_dict[{k}]')
    valid = r_prj['mask'] > 0
    assert_close(m_prj['alpha'].cpu()[valid], r_prj['alpha'][valid], rtol=1e-3, atol_scale=1e-4, what='prj_dict[alpha] (valid projections)')
    assert torch.equal(m_prj['alpha'].cpu()[~valid], r_prj['alpha'][~valid])                                  # ground state -15
    for k in ('sdf_values', 'alpha_values', 'colors_nr', 'hit_prob_nr', 'pixel_colors_nr'):
        assert_close(m_out[k].cpu(), r_out[k], what=f'network_rendering[{k}]')
    assert_close(m_out['sdf_gradient_error'].cpu(), r_out['sdf_gradient_error'], rtol=1e-3, atol_scale=1e-3, what='eikonal error')
    assert_close(m_out['s'].cpu(), r_out['s'], what='s')


def test_render_impl_matches_the_reference():
    """NeuralRayRenderer.render_impl (renderer.py:152-162), eval: coarse pass exact-path comparison; the fine pass depends on
    the ill-conditioned inverse-CDF sampler, so its colours get a loose bound (see tests/test_gpu_render.py)."""
    scn = make_scene(seed=18, num_views=4, h=96, w=160, radius=0.5)
    q = make_query(scn, 20, 6)
    ref_info = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in scn.items()}
    que = {k: torch.from_numpy(q[k]) for k in ('coords', 'poses', 'Ks', 'depth_range', 'imgs')}
    with RH.shims():
        _, ref_net = RH.build_reference_net(0)
        r = {k: v.detach() for k, v in ref_net.nr_net.render_impl(dict(que), dict(ref_info), False).items()}
    net = seed0_model().to(DEV).eval()
    with torch.no_grad():
        m = net.nr_net.render_impl({k: v.to(DEV) for k, v in que.items()}, {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in ref_info.items()}, False)
    assert set(r) <= set(m), sorted(set(r) - set(m))
    for k in ('sdf_values', 'alpha_values', 'hit_prob_nr', 'colors_nr', 'pixel_colors_nr', 'render_depth', 'pixel_colors_gt'):
        assert_close(m[k].cpu(), r[k], what=f'render_impl[{k}]')
    assert torch.equal(m['ray_mask'].cpu(), r['ray_mask'])
    assert float((m['pixel_colors_nr_fine'].cpu() - r['pixel_colors_nr_fine']).abs().max()) < 5e-3
