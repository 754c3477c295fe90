"""GPU: the backward (training) kernels of the volume path against torch autograd through the oracle (CPU) and against
the gradients of the unmodified reference (tests/golden/volume_grad_small_v4.npz; fp64 truth in volume_grad64_small_v4.npz).

Tolerance: the recipe of SURVEY.md 8c (1e-4, relative to the tensor's max) for every well-conditioned tensor.  The
compute_prob chain (dist_decoder.*, prob_embed.0, d ray_feats) is ill-conditioned in fp32: the REFERENCE's own fp32
gradients deviate from the fp64 truth by up to 3.5e-3 there (ReLU gates of prob_embed.0 sitting at the switching point,
cancellation between the near/far logistic CDFs); entry by entry our values coincide with the reference's fp32 values
except where one of the two flips a gate.  Those tensors are bounded against the fp64 truth (max 1e-2, rel-L2 5e-3)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import load_golden, golden_weights, assert_close
from tests.golden.cases import VOLUME_CASES
from graspnerf_b200.synth import make_scene

pytestmark = pytest.mark.gpu
R = 40


def _setup():
    from graspnerf_b200 import ops
    dev = torch.device('cuda:0')
    sd = {k: v for k, v in golden_weights().items() if k.startswith(('agg_net.', 'dist_decoder.'))}
    sc = make_scene(**VOLUME_CASES['small_v4'])
    sct = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in sc.items()}
    hw = ops.HeadWeights(sd, 'agg_net.', 'dist_decoder.', dev)
    scene = ops.Scene(*[sct[k].to(dev) for k in ('imgs', 'img_feats', 'ray_feats', 'poses', 'Ks', 'depth_range')])
    bbox = torch.tensor([sc['bbox3d'][0]], device=dev)
    return ops, dev, sd, sc, sct, hw, scene, bbox


def _maxrel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_k1_backward_matches_autograd_of_the_gathers():
    from oracle import nr_oracle as O
    ops, dev, sd, sc, sct, hw, scene, bbox = _setup()
    d_rec = torch.from_numpy(np.random.default_rng(5).standard_normal((1, R ** 3, scene.V, 64)).astype(np.float32)).to(dev)
    d_img, d_ray = ops.k1_backward(scene, hw, d_rec, resolution=R, bbox_min=bbox)
    imc = sct['img_feats'].clone().requires_grad_(True)
    rac = sct['ray_feats'].clone().requires_grad_(True)
    sc2 = dict(sct, img_feats=imc, ray_feats=rac)
    pts = O.volume_query_points(sc['bbox3d'][0], R, 0.3, torch.float32).reshape(-1, 3)
    with torch.enable_grad():
        r = O.project_and_sample(sc2, pts)
        dr = d_rec[0].cpu()
        ((r['ray_feats'] * dr[..., :32]).sum() + (r['img_feats'] * dr[..., 32:]).sum()).backward()
    assert_close(d_ray[0].permute(0, 3, 1, 2).cpu(), rac.grad, what='d_ray_feats')
    assert_close(d_img[0].permute(0, 3, 1, 2).cpu(), imc.grad, what='d_img_feats')


def test_k2b_backward_matches_autograd_of_the_ray_head():
    from oracle import nr_oracle as O
    from graspnerf_b200.weights import unpack_blob_grad
    ops, dev, sd, sc, sct, hw, scene, bbox = _setup()
    G = torch.from_numpy(load_golden('volume_grad_small_v4.npz')['G'])
    rec, pt = ops.k1_forward(scene, hw, resolution=R, bbox_min=bbox)
    pooled, _, _ = ops.k2a_forward(rec, pt, hw, scene.depth_range, impl='simt')
    d_w = torch.zeros(hw.blob.shape, dtype=torch.float64, device=dev)
    d_pooled = ops.k2b_backward(pooled, hw, G.to(dev), d_w, dn=R, resolution=R, bbox_min=bbox)
    A = 'agg_net.agg_impl.'
    sdc = {k: v.clone().requires_grad_(True) for k, v in sd.items()
           if k.startswith((A + 'geometry_fc', A + 'ray_attention', A + 'out_geometry'))}
    pc = pooled[0].cpu()
    pin = pc[:, :65].clone().requires_grad_(True)
    nvalid = pc[:, 65].reshape(R * R, R)
    pts = O.volume_query_points(sc['bbox3d'][0], R, 0.3, torch.float32).reshape(R * R, R, 3)
    with torch.enable_grad():                                   # ibrnet.py:485-495 from the pooled features
        g = torch.cat([pin.reshape(R * R, R, 65), O.embed_points(pts)], -1)
        g = F.elu(O._lin(sdc, A + 'geometry_fc.2', F.elu(O._lin(sdc, A + 'geometry_fc.0', g)))) + O.positional_table(R)[None]
        g = O.ray_attention(sdc, A, g, (nvalid > 1).float())
        sdf = O._lin(sdc, A + 'out_geometry_fc.1', O._lin(sdc, A + 'out_geometry_fc.0', g)).clip(-1, 1)[..., 0]
        sdf = sdf.masked_fill(nvalid < 1, 1.0)
        (sdf.reshape(1, 1, R, R, R).flip(-1) * G).sum().backward()
    assert_close(d_pooled[0, :, :65].cpu(), pin.grad, what='d_pooled')
    gk = unpack_blob_grad(d_w.float().cpu())
    for k in sdc:
        assert_close(gk[k], sdc[k].grad, what=k)


def test_volume_path_gradients_end_to_end():
    ops, dev, sd, sc, sct, hw, scene, bbox = _setup()
    g32 = load_golden('volume_grad_small_v4.npz')              # unmodified reference, fp32 autograd
    g64 = load_golden('volume_grad64_small_v4.npz')            # oracle in fp64
    params = {k: torch.nn.Parameter(v.clone().to(dev)) for k, v in sd.items()}
    imgf = sct['img_feats'].to(dev).requires_grad_(True)
    rayf = sct['ray_feats'].to(dev).requires_grad_(True)
    vol = ops.sample_volume_autograd(sct['imgs'].to(dev), imgf, rayf, sct['poses'].to(dev), sct['Ks'].to(dev),
                                     sct['depth_range'].to(dev), bbox, params, R)
    loss = (vol * torch.from_numpy(g32['G']).to(dev)).sum()
    loss.backward()
    assert abs(loss.item() - float(g32['loss'])) <= 1e-4 * abs(float(g32['loss']))
    got = {'d_img_feats': imgf.grad.cpu().numpy(), 'd_ray_feats': rayf.grad.cpu().numpy()}
    for k, p in params.items():
        if 'rgb_fc' in k or 'deviation' in k:
            assert p.grad is None, k                            # not on the volume path
        else:
            assert p.grad is not None, k
            got['dw/' + k] = p.grad.cpu().numpy()
    def l2rel(a, b):
        return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))
    report = []
    for k, a in got.items():
        ill = k.startswith(('dw/dist_decoder.', 'dw/agg_net.prob_embed.0')) or k == 'd_ray_feats'
        if ill:
            # compute_prob chain / ReLU gates of prob_embed.0: single gate flips between two fp32 evaluations move an entry by
            # ~1e-2 of the tensor's max (the reference's own fp32 gradient shows the same events against fp64): bound the
            # max error at 1e-2 and the rel-L2 error at 5e-3, both against the fp64 truth
            ok = _maxrel(a, g64[k]) <= 1e-2 and l2rel(a, g64[k]) <= 5e-3
        elif 'neuray_fc' in k:
            ok = _maxrel(a, g64[k]) <= 5e-4
        else:                                                   # well-conditioned part: the 1e-4 recipe against the REFERENCE's fp32 gradients
            ok = _maxrel(a, g32[k]) <= 1e-4
        report.append((k, ok, _maxrel(a, g64[k]), _maxrel(g32[k], g64[k])))
    bad = [r for r in report if not r[1]]
    assert not bad, 'gradient outside tolerance (key, ok, ours-vs-fp64, ref32-vs-fp64): ' + repr(bad)


def test_mirror_trains_through_the_cuda_backward():
    """GraspNeRF.forward under autograd (render_rgb off): a loss on volume + vgn heads reaches the 2-D encoders and the
    head weights through the hand-written backward kernels; one Adam step changes the volume."""
    from graspnerf_b200.network import name2network
    from graspnerf_b200.synth import make_query
    from tests.test_boundary import CFG
    dev = torch.device('cuda:0')
    cfg = dict(CFG, render_rgb=False)
    torch.manual_seed(0)
    net = name2network[cfg['network']](cfg).to(dev).train()
    scene = make_scene(**VOLUME_CASES['small_v4'])
    ref = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in scene.items()
           if k not in ('img_feats', 'ray_feats')}
    q = {k: torch.from_numpy(v).to(dev) for k, v in make_query(scene, 16, 7).items() if isinstance(v, np.ndarray)}
    data = {'step': 0, 'full_vol': True, 'ref_imgs_info': ref, 'que_imgs_info': q, 'src_imgs_info': ref}
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    sdf_gt = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (1, 1, R, R, R)).astype(np.float32)).to(dev)
    out = net(data)
    loss = F.smooth_l1_loss(out['volume'], sdf_gt) + out['vgn_pred'][0].mean() * 0.1
    loss.backward()
    named = dict(net.named_parameters())
    for k in ('nr_net.agg_net.agg_impl.base_fc.0.weight', 'nr_net.dist_decoder.mean_decoder.0.weight',
              'nr_net.image_encoder.conv1.weight', 'nr_net.agg_net.agg_impl.ray_attention.w_qs.weight'):
        assert named[k].grad is not None and torch.isfinite(named[k].grad).all() and named[k].grad.abs().sum() > 0, k
    v0 = out['volume'].detach().clone()
    opt.step()
    with torch.no_grad():
        v1 = net(data)['volume']
    assert (v1 - v0).abs().max() > 0


def test_batched_backward_equals_per_scene_backward():
    """B = 2 scenes in one launch give each scene's feature-map gradients and the SUM of the two weight gradients."""
    from graspnerf_b200 import ops
    dev = torch.device('cuda:0')
    sd = {k: v for k, v in golden_weights().items() if k.startswith(('agg_net.', 'dist_decoder.'))}
    hw = ops.HeadWeights(sd, 'agg_net.', 'dist_decoder.', dev)
    scs = [make_scene(seed=s, num_views=4, h=96, w=160, radius=0.45) for s in (3, 4)]
    G = torch.from_numpy(np.random.default_rng(2).standard_normal((2, 1, R, R, R)).astype(np.float32)).to(dev)

    def run(idx):
        st = lambda k: torch.from_numpy(np.stack([scs[i][k] for i in idx])).to(dev)
        scene = ops.Scene(st('imgs'), st('img_feats'), st('ray_feats'), st('poses'), st('Ks'), st('depth_range'))
        bbox = torch.tensor([scs[i]['bbox3d'][0] for i in idx], device=dev)
        rec, pt = ops.k1_forward(scene, hw, resolution=R, bbox_min=bbox)
        pooled, _, _ = ops.k2a_forward(rec, pt, hw, scene.depth_range)
        return ops.sample_volume_backward(scene, hw, bbox, rec, pt, pooled, G[idx].contiguous(), R)
    di, dr, dw = run([0, 1])
    di0, dr0, dw0 = run([0])
    di1, dr1, dw1 = run([1])
    assert_close(di[0].cpu(), di0[0].cpu(), what='d_img scene 0'); assert_close(di[1].cpu(), di1[0].cpu(), what='d_img scene 1')
    assert_close(dr[0].cpu(), dr0[0].cpu(), rtol=1e-3, atol_scale=1e-3, what='d_ray scene 0')
    assert_close(dr[1].cpu(), dr1[0].cpu(), rtol=1e-3, atol_scale=1e-3, what='d_ray scene 1')
    assert_close(dw.cpu(), (dw0 + dw1).cpu(), rtol=1e-3, atol_scale=1e-3, what='weight gradient blob')
