"""Generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the authoring container only:  python tests/golden/make_golden.py
Inputs are regenerated from seeds by graspnerf_b200.synth (numpy PCG64), so only the
reference's weights and OUTPUTS are stored.  Reference call sites exercised:
  NeuralRayRenderer.sample_volume  (renderer.py:164-199)  -> volume + stage intermediates
  NeuralRayRenderer.render_impl    (renderer.py:152-162)  -> coarse+fine RGB head (eval)
"""
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ref_harness import build_reference_net            # noqa: E402
from graspnerf_b200.synth import make_scene, make_query  # noqa: E402

HOT_PREFIXES = ('agg_net.', 'fine_agg_net.', 'dist_decoder.', 'fine_dist_decoder.')

from cases import VOLUME_CASES, RENDER_CASES  # noqa: E402


def to_torch(d):
    return {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in d.items()}


def hot_state_dict(nr):
    return {k: v.detach().clone() for k, v in nr.state_dict().items() if k.startswith(HOT_PREFIXES)}


def run_volume_case(nr, kw, n_probe=512):
    from network.render_ops import project_points_dict
    from utils.field_utils import TSDF_SAMPLE_POINTS
    scene = make_scene(**kw)
    ref = to_torch(scene)
    out = {}
    with torch.no_grad():
        vol = nr.sample_volume(ref)
        out['volume'] = vol[0, 0].numpy()
        # re-run the stages one by one (same calls sample_volume makes) to record intermediates
        res = 40
        que_pts = (torch.from_numpy(TSDF_SAMPLE_POINTS) + torch.tensor(ref['bbox3d'][0])).reshape(1, res * res, res, 3)
        que_pts = torch.flip(que_pts, (2,))
        prj = project_points_dict(ref, que_pts)
        prj = nr.get_img_feats(ref, prj)
        prj = nr.predict_proj_ray_prob(prj, ref, torch.empty(0), False)
        que_dir = torch.tensor([0, 0, 1]).reshape(1, 1, 1, 3).repeat(1, res * res, res, 1)
        feats, emb, dir_diff, vmask = nr.agg_net._get_embedding(prj, que_dir)
        outs, _ = nr.agg_net.agg_impl(feats, emb, dir_diff, vmask, que_pts)
        assert torch.equal(outs[..., 3].reshape(1, 1, res, res, res).flip(-1), vol)
    V = ref['imgs'].shape[0]
    N = res ** 3

    def nv(t):  # [V,1,rn,dn,C] -> [N,V,C]
        return t.reshape(V, N, -1).permute(1, 0, 2)
    rng = np.random.default_rng(1234)
    probe = np.sort(rng.choice(N, n_probe, replace=False))
    out['mask'] = nv(prj['mask'])[..., 0].numpy().astype(np.uint8)            # full [N,V] index table
    out['probe'] = probe.astype(np.int64)
    for k in ('pts', 'depth', 'dir', 'rgb', 'ray_feats', 'img_feats', 'hit_prob', 'vis', 'alpha'):
        out['p_' + ('uv' if k == 'pts' else k)] = nv(prj[k])[probe].numpy()
    out['p_prob_emb'] = emb.reshape(N, V, -1)[probe].numpy()
    out['p_dir_diff'] = dir_diff.reshape(N, V, -1)[probe].numpy()
    out['valid_ratio'] = np.float32(prj['mask'].mean().item())
    return out


def run_render_case(nr, case):
    scene = make_scene(**case['scene'])
    ref = to_torch(scene)
    q = make_query(scene, case['num_rays'], case['qseed'])
    que = to_torch(q)
    # torch.autograd.grad inside the reference needs grad mode on; eval = is_train False
    res = nr.render_impl(que, ref, False)
    keep = ['pixel_colors_nr', 'alpha_values', 'hit_prob_nr', 'sdf_values', 'colors_nr', 'render_depth',
            'ray_mask', 'sdf_gradient_error', 'pixel_colors_gt', 's']
    out = {}
    for k in keep:
        for sfx in ('', '_fine'):
            if k + sfx in res:
                out[k + sfx] = res[k + sfx].detach().numpy()
    # the fine pass's own inputs, recomputed with the reference's samplers (eval => deterministic)
    from network.render_ops import sample_depth, sample_fine_depth
    with torch.no_grad():
        depth, _ = sample_depth(que['depth_range'], que['coords'], nr.cfg['depth_sample_num'], False)
        fd = sample_fine_depth(depth, res['hit_prob_nr'].detach(), que['depth_range'],
                               nr.cfg['fine_depth_sample_num'], False)
        out['depth'] = depth.numpy()
        out['fine_depth_unsorted'] = fd.numpy()
        out['depth_fine'] = torch.sort(fd, -1)[0].numpy()
    return out


def main():
    cfg, net = build_reference_net(0)
    nr = net.nr_net
    sd = hot_state_dict(nr)
    np.savez_compressed(os.path.join(HERE, 'weights_seed0.npz'), **{k: v.numpy() for k, v in sd.items()})
    print('weights', len(sd), sum(v.numel() for v in sd.values()))
    for name, kw in VOLUME_CASES.items():
        out = run_volume_case(nr, kw)
        np.savez_compressed(os.path.join(HERE, f'volume_{name}.npz'), **out)
        print(name, 'valid_ratio', out['valid_ratio'], 'vol mean/std', out['volume'].mean(), out['volume'].std())
    for name, case in RENDER_CASES.items():
        out = run_render_case(nr, case)
        np.savez_compressed(os.path.join(HERE, f'render_{name}.npz'), **out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == '__main__' and len(sys.argv) == 1:
    main()


def make_model_golden():
    """state_dict key/shape/checksum table of the reference model under torch.manual_seed(0) and one full
    GraspNeRF.forward (eval, render_rgb off as main.py:150 does) on a small scene."""
    import json
    cfg, net = build_reference_net(0)
    table = {k: [list(v.shape), float(v.double().sum()), float(v.double().abs().sum())] for k, v in net.state_dict().items()}
    with open(os.path.join(HERE, 'state_dict_keys.json'), 'w') as f:
        json.dump(table, f, indent=0)
    kw = dict(seed=3, num_views=4, h=96, w=160, radius=0.45)
    scene = make_scene(**kw)
    ref = to_torch({k: v for k, v in scene.items() if k not in ('img_feats', 'ray_feats')})
    q = to_torch(make_query(scene, 16, 7))
    net.nr_net.cfg['render_rgb'] = False
    data = {'step': 0, 'eval': True, 'full_vol': True, 'ref_imgs_info': ref, 'que_imgs_info': q,
            'src_imgs_info': ref}
    torch.manual_seed(123)
    with torch.no_grad():
        out = net(data)
    np.savez_compressed(os.path.join(HERE, 'forward_small_v4.npz'), volume=out['volume'][0, 0].numpy(),
                        qual=out['vgn_pred'][0][0, 0].numpy(), depth_mean=out['depth_mean'].numpy(),
                        depth_coords=out['depth_coords'].numpy())
    print('model golden: keys', len(table), 'volume std', out['volume'].std().item(), sorted(out.keys()))


if __name__ == '__main__' and '--model' in sys.argv:
    make_model_golden()


def make_grad_golden():
    """First-order gradients of the volume path from the UNMODIFIED reference (CPU autograd):
    loss = sum(volume * G), G ~ N(0,1) seeded; d loss / d {img_feats, ray_feats, every agg_net.* / dist_decoder.* parameter}.
    Stored for the small_v4 case (tests/golden/volume_grad_small_v4.npz)."""
    cfg, net = build_reference_net(0)
    nr = net.nr_net
    kw = VOLUME_CASES['small_v4']
    ref = to_torch(make_scene(**kw))
    ref['img_feats'].requires_grad_(True)
    ref['ray_feats'].requires_grad_(True)
    G = torch.from_numpy(np.random.default_rng(99).standard_normal((1, 1, 40, 40, 40)).astype(np.float32))
    vol = nr.sample_volume(ref)
    loss = (vol * G).sum()
    loss.backward()
    out = {'G': G.numpy(), 'loss': np.float64(loss.item()), 'd_img_feats': ref['img_feats'].grad.numpy(),
           'd_ray_feats': ref['ray_feats'].grad.numpy()}
    n = 0
    for k, v in nr.named_parameters():
        if k.startswith(('agg_net.', 'dist_decoder.')) and v.grad is not None:
            out['dw/' + k] = v.grad.numpy(); n += 1
    np.savez_compressed(os.path.join(HERE, 'volume_grad_small_v4.npz'), **out)
    print('grad golden:', n, 'parameter grads; |d_img_feats|', float(ref['img_feats'].grad.abs().sum()),
          '|d_ray_feats|', float(ref['ray_feats'].grad.abs().sum()))


if __name__ == '__main__' and '--grads' in sys.argv:
    make_grad_golden()


RENDER_GRAD_CASE = dict(scene=dict(seed=3, num_views=4, h=96, w=160, radius=0.45), num_rays=32, qseed=7, useed=11, gseed=77)


def make_render_grad_golden():
    """TRAINING-mode gradients of the RGB head from the UNMODIFIED reference (CPU autograd, is_train=True):
    loss = sum(pixel_colors_nr * G1) + sum(pixel_colors_nr_fine * G2) + 0.1 * (sdf_gradient_error + sdf_gradient_error_fine)
    (the eikonal terms need the second derivative of the per-ray head, ibrnet.py:497-504).  torch.rand inside
    sample_fine_depth (render_ops.py:205) is replaced by a seeded u that is stored with the fixture.
    -> tests/golden/render_grad_small_v4.npz"""
    cfg, net = build_reference_net(0)
    nr = net.nr_net
    case = RENDER_GRAD_CASE
    scene = make_scene(**case['scene'])
    ref = to_torch(scene)
    ref['img_feats'].requires_grad_(True)
    ref['ray_feats'].requires_grad_(True)
    que = to_torch(make_query(scene, case['num_rays'], case['qseed']))
    rn, fdn = case['num_rays'], nr.cfg['fine_depth_sample_num']
    u = torch.from_numpy(np.random.default_rng(case['useed']).random((1, rn, fdn), dtype=np.float32))
    rng = np.random.default_rng(case['gseed'])
    G1 = torch.from_numpy(rng.standard_normal((1, rn, 3)).astype(np.float32))
    G2 = torch.from_numpy(rng.standard_normal((1, rn, 3)).astype(np.float32))
    real_rand = torch.rand
    torch.rand = lambda *a, **k: u.clone()
    try:
        res = nr.render_impl(que, ref, True)
    finally:
        torch.rand = real_rand
    loss = (res['pixel_colors_nr'] * G1).sum() + (res['pixel_colors_nr_fine'] * G2).sum() \
        + 0.1 * (res['sdf_gradient_error'].sum() + res['sdf_gradient_error_fine'].sum())
    loss.backward()
    # the fine pass's own sample depths (the inverse-CDF sampler is ill-conditioned where the coarse hit probability is ~0:
    # parity tests feed THESE depths to both sides, like the eval fixtures do)
    from network.render_ops import sample_depth, sample_fine_depth
    with torch.no_grad():
        depth, _ = sample_depth(que['depth_range'], que['coords'], nr.cfg['depth_sample_num'], False)
        torch.rand = lambda *a, **k: u.clone()
        try:
            fd = sample_fine_depth(depth, res['hit_prob_nr'].detach(), que['depth_range'], fdn, True)
        finally:
            torch.rand = real_rand
    out = {'depth_fine': torch.sort(fd, -1)[0].numpy(), 'u': u.numpy(), 'G1': G1.numpy(), 'G2': G2.numpy(), 'loss': np.float64(loss.item()),
           'd_img_feats': ref['img_feats'].grad.numpy(), 'd_ray_feats': ref['ray_feats'].grad.numpy()}
    for k in ('pixel_colors_nr', 'pixel_colors_nr_fine', 'sdf_gradient_error', 'sdf_gradient_error_fine', 'alpha_values',
              'alpha_values_fine', 'hit_prob_nr', 'sdf_values', 'sdf_values_fine', 'render_depth', 'render_depth_fine'):
        out[k] = res[k].detach().numpy()
    n = 0
    for k, v in nr.named_parameters():
        if k.startswith(HOT_PREFIXES) and v.grad is not None:
            out['dw/' + k] = v.grad.numpy(); n += 1
    np.savez_compressed(os.path.join(HERE, 'render_grad_small_v4.npz'), **out)
    print('render grad golden:', n, 'parameter grads; loss', loss.item(), '|d_img_feats|', float(ref['img_feats'].grad.abs().sum()),
          '|d_ray_feats|', float(ref['ray_feats'].grad.abs().sum()), 'eik', res['sdf_gradient_error'].item(), res['sdf_gradient_error_fine'].item())


if __name__ == '__main__' and '--render-grads' in sys.argv:
    make_render_grad_golden()


def loss_case(seed=5, rn=24, pn=64, rfn=3, h=20, w=28, G=16, R=8):
    """Random (data_pr, data_gt) with the keys the four losses of nrvgn_sdf.yaml read; regenerated from the seed by the
    test, so the fixture stores only the reference's loss values."""
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
    sdf_gt[0, 0, :3] = -1.0                                   # masked voxels (loss.py:169)
    data_gt = {'scene_name': 'vgn_syn/0', 'ref_imgs_info': {'true_depth': f(rfn, 1, h, w) * 0.6 + 0.2,
               'depth_range': torch.tensor([[0.2, 0.8]] * rfn), 'sdf_gt': sdf_gt},
               'grasp_info': [torch.from_numpy(rng.integers(0, R, (G, 3))), torch.from_numpy((rng.random(G) < 0.5).astype(np.float32)),
                              torch.from_numpy(quat), f(G) * 10]}
    return data_pr, data_gt


def make_loss_golden():
    """Values of the reference's own loss classes (network/loss.py: RenderLoss, DepthLoss, SDFLoss, VGNLoss with the shipped
    yaml) on loss_case() -> tests/golden/losses.json.  pyquaternion / torchmetrics are stubbed (imported, unused here)."""
    import json, types
    class _Stub(types.ModuleType):                       # missing third-party modules: importable, every attribute is a stub
        def __getattr__(self, k):
            if k.startswith('__'):
                raise AttributeError(k)
            return _Stub(self.__name__ + '.' + k)

        def __call__(self, *a, **k):
            return self
    import importlib.abc, importlib.machinery

    class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        ROOTS = ('pyquaternion', 'torchmetrics', 'skimage', 'transforms3d', 'plyfile', 'h5py', 'open3d', 'inplace_abn', 'kornia', 'lpips')

        def find_spec(self, name, path, target=None):
            if name.split('.')[0] in self.ROOTS:
                return importlib.machinery.ModuleSpec(name, self, is_package=True)
            return None

        def create_module(self, spec):
            return _Stub(spec.name)

        def exec_module(self, module):
            module.__path__ = []
    sys.meta_path.append(_Finder())
    cfg, _ = __import__('ref_harness').load_reference()
    from network.loss import name2loss
    data_pr, data_gt = loss_case()
    out = {}
    for name in cfg['loss']:
        res = name2loss[name](cfg)(data_pr, data_gt, 0, is_train=True)
        for k, v in res.items():
            out[k] = float(torch.as_tensor(v).reshape(-1)[0])
    with open(os.path.join(HERE, 'losses.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print('loss golden:', out)


if __name__ == '__main__' and '--losses' in sys.argv:
    make_loss_golden()


def quantise_images(scene):
    """uint8 images as the planner holds them (main.py:166-171) and their fp32 form u8 / 255 (color_map_forward)."""
    u8 = np.clip(np.floor(scene['imgs'] * 256.0), 0, 255).astype(np.uint8)
    scene = dict(scene)
    scene['imgs'] = u8.astype(np.float32) / np.float32(255.0)
    return scene, np.ascontiguousarray(u8.transpose(0, 2, 3, 1))


def make_extra_golden():
    """Round-2 fixtures (python tests/golden/make_golden.py --extra):
      render_inds.npz        the int64 `inds` table torch.searchsorted returns INSIDE the reference's sample_fine_depth
                             (render_ops.py:210) for every render case - the reference-held index table K3 is compared with;
      forward_small_u8.npz   one full GraspNeRF.forward (eval, render_rgb off, main.py:150) on a small scene whose images are
                             uint8 / 255 (what the planner feeds): volume, qual, rot, width, depth_mean (+ coords)."""
    cfg, net = build_reference_net(0)
    from network.render_ops import sample_depth, sample_fine_depth
    nr = net.nr_net
    inds_out = {}
    real_ss = torch.searchsorted
    for name, case in RENDER_CASES.items():
        scene = make_scene(**case['scene'])
        ref, que = to_torch(scene), to_torch(make_query(scene, case['num_rays'], case['qseed']))
        res = nr.render_impl(que, ref, False)
        seen = []

        def spy(*a, **k):
            r = real_ss(*a, **k)
            seen.append(r.clone())
            return r
        torch.searchsorted = spy
        try:
            with torch.no_grad():
                depth, _ = sample_depth(que['depth_range'], que['coords'], nr.cfg['depth_sample_num'], False)
                sample_fine_depth(depth, res['hit_prob_nr'].detach(), que['depth_range'], nr.cfg['fine_depth_sample_num'], False)
        finally:
            torch.searchsorted = real_ss
        assert len(seen) == 1
        inds_out[name] = seen[0].numpy().astype(np.int64)
        print('inds', name, inds_out[name].shape, inds_out[name].min(), inds_out[name].max())
    np.savez_compressed(os.path.join(HERE, 'render_inds.npz'), **inds_out)

    scene, u8 = quantise_images(make_scene(seed=3, num_views=4, h=96, w=160, radius=0.45))
    ref = to_torch({k: v for k, v in scene.items() if k not in ('img_feats', 'ray_feats')})
    q = to_torch(make_query(scene, 16, 7))
    net.nr_net.cfg['render_rgb'] = False
    data = {'step': 0, 'eval': True, 'full_vol': True, 'ref_imgs_info': ref, 'que_imgs_info': q, 'src_imgs_info': ref}
    torch.manual_seed(123)
    with torch.no_grad():
        out = net(data)
    qual, rot, width = out['vgn_pred']
    np.savez_compressed(os.path.join(HERE, 'forward_small_u8.npz'), volume=out['volume'][0, 0].numpy(), qual=qual[0, 0].numpy(),
                        rot=rot[0].numpy(), width=width[0, 0].numpy(), depth_mean=out['depth_mean'].numpy(),
                        depth_coords=out['depth_coords'].numpy(), imgs_u8_checksum=np.int64(u8.astype(np.int64).sum()))
    print('forward_small_u8: volume std', out['volume'].std().item(), 'qual range', qual.min().item(), qual.max().item())


if __name__ == '__main__' and '--extra' in sys.argv:
    make_extra_golden()
