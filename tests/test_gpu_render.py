"""GPU parity of the RGB head (K1 ray mode -> K2a(+rgb_fc) -> K2b(+grad) -> K3) against the oracle and the reference's
committed outputs (tests/golden/render_*.npz).  Index tables (coarse depths, ray_mask, searchsorted inds) bit-exact."""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden, golden_weights, assert_close
from tests.golden.cases import RENDER_CASES
from graspnerf_b200.synth import make_scene, make_query

pytestmark = pytest.mark.gpu


def _setup(name):
    from graspnerf_b200 import ops
    case = RENDER_CASES[name]
    sd = golden_weights()
    scn = make_scene(**case['scene'])
    sc = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in scn.items()}
    q = make_query(scn, case['num_rays'], case['qseed'])
    dev = torch.device('cuda:0')
    scene = ops.Scene(*[sc[k].to(dev) for k in ('imgs', 'img_feats', 'ray_feats', 'poses', 'Ks', 'depth_range')])
    que = {k: torch.from_numpy(q[k]).to(dev) for k in ('coords', 'poses', 'Ks', 'depth_range')}
    oq = {'coords': torch.from_numpy(q['coords'][0]), 'pose': torch.from_numpy(q['poses'][0]),
          'K': torch.from_numpy(q['Ks'][0]), 'depth_range': torch.from_numpy(q['depth_range'][0])}
    hw_c = ops.HeadWeights(sd, 'agg_net.', 'dist_decoder.', dev)
    hw_f = ops.HeadWeights(sd, 'fine_agg_net.', 'fine_dist_decoder.', dev)
    return ops, sd, sc, scene, que, oq, hw_c, hw_f, dev


@pytest.mark.parametrize('impl', ['simt', 'tc'])
@pytest.mark.parametrize('name', list(RENDER_CASES))
def test_render_coarse_and_fine(name, impl):
    from oracle import nr_oracle as O
    ops, sd, sc, scene, que, oq, hw_c, hw_f, dev = _setup(name)
    ops.K2A_IMPL = impl
    try:
        g = load_golden(f'render_{name}.npz')
        keys = ('sdf_values', 'alpha_values', 'hit_prob_nr', 'colors_nr', 'pixel_colors_nr', 'render_depth')
        # ---- coarse pass
        depth = ops.k3_coarse_depths(que['depth_range'], que['coords'].shape[1], 40)
        assert np.array_equal(depth.cpu().numpy(), g['depth']), 'coarse depth table differs'
        out = ops.render_by_depth(scene, hw_c, que, depth)
        oc = O.render_by_depth(sd, sc, oq, torch.from_numpy(g['depth'][0]), False)
        assert_close(out['sdf_grad'][0].cpu(), oc['sdf_grad'], rtol=1e-3, atol_scale=1e-3, what='d sdf / d pts vs oracle autograd')
        assert np.array_equal(out['ray_mask'].cpu().numpy(), g['ray_mask'])
        for k in keys:
            assert_close(out[k][0].cpu(), oc[k], what=f'{k} vs oracle')
            assert_close(out[k].cpu(), g[k], what=f'{k} vs reference golden')
        assert_close(out['sdf_gradient_error'].cpu(), g['sdf_gradient_error'], rtol=1e-3, atol_scale=1e-3, what='eikonal')
        # ---- hierarchical sampler on the reference's coarse hit_prob: bit-exact vs the oracle's fixed-order restatement
        hp = torch.from_numpy(g['hit_prob_nr']).to(dev)
        u = (0.5 / 40 + torch.arange(40, dtype=torch.float32) / 40).expand(1, hp.shape[1], 40).contiguous()
        fd, inds = ops.k3_fine_depths(torch.from_numpy(g['depth']).to(dev), hp, que['depth_range'], u.to(dev), want_inds=True)
        ofd, oinds = O.fine_depths(torch.from_numpy(g['depth'][0]), torch.from_numpy(g['hit_prob_nr'][0]), oq['depth_range'], 40)
        assert torch.equal(inds[0].cpu(), oinds), 'searchsorted index table differs from the oracle'
        # ... and vs the table torch.searchsorted returned INSIDE the unmodified reference's sample_fine_depth (render_ops.py:210)
        ref_inds = load_golden('render_inds.npz')[name]
        assert np.array_equal(inds.cpu().numpy(), ref_inds), 'searchsorted index table differs from the reference-held table'
        assert_close(fd[0].cpu(), torch.sort(ofd, -1)[0], rtol=1e-5, atol_scale=1e-6, what='fine depths vs oracle')
        assert_close(fd.cpu(), g['depth_fine'], what='fine depths vs reference')
        # ---- fine pass on the reference's fine depths
        fine = ops.render_by_depth(scene, hw_f, que, torch.from_numpy(g['depth_fine']).to(dev))
        assert np.array_equal(fine['ray_mask'].cpu().numpy(), g['ray_mask_fine'])
        for k in keys:
            assert_close(fine[k].cpu(), g[k + '_fine'], what=f'{k}_fine vs reference golden')
        # ---- the chain's OWN fine pass (own coarse hit_prob -> own sampler -> fine kernels): the inverse-CDF sampler is
        # ill-conditioned where the coarse hit probability vanishes, so an index may legitimately flip; report, bound loosely
        full = ops.render_rays(scene, hw_c, hw_f, que, 40, 40)
        flips = int((full['fine_inds'].cpu().numpy() != ref_inds).sum())
        dmax = float((full['depth_fine'].cpu() - torch.from_numpy(g['depth_fine'])).abs().max())
        print(f'[{name}/{impl}] own-chain fine pass: {flips}/{ref_inds.size} searchsorted flips vs the reference, max |d depth| {dmax:.2e}')
        pc = full['pixel_colors_nr_fine'].cpu().numpy()
        print(f'[{name}/{impl}] own-chain fine colours: max |diff| vs the reference {np.abs(pc - g["pixel_colors_nr_fine"]).max():.2e}')
        assert np.isfinite(pc).all() and flips <= ref_inds.size // 20          # reported, not a parity claim (measured: <= 1e-2 colour diff)
    finally:
        ops.K2A_IMPL = 'tc'


@pytest.mark.parametrize('impl', ['simt', 'tc'])
def test_ragged_ray_batch(impl):
    """37 rays x 11 samples = 407 points: not a multiple of any tile size on the path (K1 32 / 16 points, K2a 16-20 points per
    128-row tile, K2b 11 rays per CTA), sample count != 40 (positional table, attention length) - against the oracle."""
    from oracle import nr_oracle as O
    ops, sd, sc, scene, que, oq, hw_c, hw_f, dev = _setup('rays48_small')
    ops.K2A_IMPL = impl
    try:
        rn, dn = 37, 11
        que = {k: (v[:, :rn].contiguous() if k == 'coords' else v) for k, v in que.items()}
        oq = dict(oq, coords=oq['coords'][:rn])
        depth = ops.k3_coarse_depths(que['depth_range'], rn, dn)
        out = ops.render_by_depth(scene, hw_c, que, depth)
        oc = O.render_by_depth(sd, sc, oq, depth[0].cpu(), False)
        for k in ('sdf_values', 'alpha_values', 'hit_prob_nr', 'colors_nr', 'pixel_colors_nr', 'render_depth'):
            assert_close(out[k][0].cpu(), oc[k], what=f'{k} vs oracle (ragged)')
        assert torch.equal(out['ray_mask'][0].cpu(), oc['ray_mask'])
    finally:
        ops.K2A_IMPL = 'tc'
