"""GPU parity of the RGB head in TRAINING: ops.ray_features_autograd (gn_k1_forward / gn_k2a_forward_tc, gn_k2a_backward incl.
rgb_fc + softmax blend, gn_k1_backward in ray mode) + the torch per-ray head / compositing (network/ray_head.py) against
 (a) torch autograd through the oracle on the same inputs and (b) the gradients of the UNMODIFIED reference
     (tests/golden/render_grad_small_v4.npz; loss incl. both eikonal terms, i.e. the second derivative of the per-ray head).
Tolerances as in tests/test_gpu_backward.py: well-conditioned tensors 2e-3 of the tensor's max (observed ~1e-5 .. 1e-4), the
fp32-ill-conditioned compute_prob chain (d ray_feats, dist_decoder.*, prob_embed.*, neuray_fc.*) 3e-2."""
import numpy as np
import pytest
import torch

from tests.helpers import (load_golden, golden_weights, assert_close, grad_close, oracle_render_grads, render_grad_loss,
                           RENDER_GRAD_CASE, HOT_PREFIXES)
from graspnerf_b200.synth import make_scene, make_query

pytestmark = pytest.mark.gpu


def _mirror_with_golden_weights(dev):
    from graspnerf_b200.network import name2network
    from tests.test_boundary import CFG
    torch.manual_seed(0)
    net = name2network[CFG['network']](dict(CFG)).to(dev).train()
    sd = golden_weights()
    named = dict(net.nr_net.named_parameters())
    with torch.no_grad():
        for k, v in sd.items():
            named[k].copy_(v.to(dev))
    return net


def test_render_training_gradients_vs_oracle_and_reference():
    from graspnerf_b200.network import ray_head
    dev = torch.device('cuda:0')
    g = load_golden('render_grad_small_v4.npz')
    net = _mirror_with_golden_weights(dev)
    nr = net.nr_net
    case = RENDER_GRAD_CASE
    scn = make_scene(**case['scene'])
    ref = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in scn.items()}
    ref['img_feats'].requires_grad_(True)
    ref['ray_feats'].requires_grad_(True)
    q = make_query(scn, case['num_rays'], case['qseed'])
    que = {k: torch.from_numpy(q[k]).to(dev) for k in ('coords', 'poses', 'Ks', 'depth_range')}
    u = torch.from_numpy(g['u']).to(dev)
    out = ray_head.render_rays_autograd(nr, ref, que, 40, 40, u, is_train=True, fine_depth=torch.from_numpy(g['depth_fine']).to(dev))
    assert nr.agg_net.step == 1 and nr.agg_net.deviation_network.variance.requires_grad        # aggregate_net.py:135-137, neus.py:17-18
    loss = render_grad_loss(out, g)
    loss.backward()
    torch.cuda.synchronize()
    # ---- forward values vs the reference
    for k in ('pixel_colors_nr', 'pixel_colors_nr_fine', 'alpha_values', 'alpha_values_fine', 'sdf_values', 'sdf_values_fine',
              'render_depth', 'render_depth_fine'):
        assert_close(out[k].detach().cpu(), g[k], what=k)
    assert_close(out['sdf_gradient_error'].detach().cpu(), g['sdf_gradient_error'], rtol=1e-3, atol_scale=1e-3, what='eikonal')
    assert abs(float(loss) - float(g['loss'])) <= 2e-4 * max(1.0, abs(float(g['loss'])))
    # ---- gradients vs the oracle's autograd (same fp32 formulas on CPU) and vs the reference
    _, _, o_img, o_ray, o_dw = oracle_render_grads(g)
    ours = {k: v.grad for k, v in nr.named_parameters() if k.startswith(HOT_PREFIXES) and v.grad is not None}
    report = []
    for name, a, b_or, b_ref, loose in (('d img_feats', ref['img_feats'].grad, o_img, g['d_img_feats'], False),
                                        ('d ray_feats', ref['ray_feats'].grad, o_ray, g['d_ray_feats'], True)):
        tol = 3e-2 if loose else 2e-3
        report.append((name,) + grad_close(a.cpu(), b_ref, tol, tol, name + ' vs reference'))
        grad_close(a.cpu(), b_or, tol, tol, name + ' vs oracle')
    keys = [k[3:] for k in g if k.startswith('dw/')]
    assert len(keys) == 126
    for k in keys:
        assert k in ours, f'no gradient reached {k}'
        loose = ('dist_decoder' in k) or ('prob_embed' in k) or ('neuray_fc' in k)
        tol = 3e-2 if loose else 2e-3
        report.append((k,) + grad_close(ours[k].cpu(), g['dw/' + k], tol, tol, k + ' vs reference'))
        grad_close(ours[k].cpu(), o_dw[k], tol, tol, k + ' vs oracle')
    worst = sorted(report, key=lambda r: -r[1])[:5]
    print('worst gradient errors vs reference (key, max-rel, rel-L2):', worst)


def test_mirror_trains_with_render_rgb_on():
    """GraspNeRF.forward with the SHIPPED configuration (render_rgb: true, nrvgn_sdf.yaml) under autograd: render loss
    (loss.py:66-84) + eikonal term (loss.py:172-173) + SDF loss reach the 2-D encoders, both head weight sets and the NeuS
    variance; one Adam step changes the rendered colours."""
    import torch.nn.functional as F
    dev = torch.device('cuda:0')
    net = _mirror_with_golden_weights(dev)
    scene = make_scene(seed=3, num_views=4, h=96, w=160, radius=0.45)
    ref = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in scene.items()
           if k not in ('img_feats', 'ray_feats')}
    q = {k: torch.from_numpy(v).to(dev) for k, v in make_query(scene, 24, 7).items() if isinstance(v, np.ndarray)}
    data = {'step': 0, 'full_vol': True, 'ref_imgs_info': ref, 'que_imgs_info': q, 'src_imgs_info': ref}
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    torch.manual_seed(1)
    out = net(data)
    for k in ('pixel_colors_nr', 'pixel_colors_nr_fine', 'pixel_colors_gt', 'ray_mask', 'sdf_gradient_error', 'volume', 'render_depth', 's'):
        assert k in out, k
    m = out['ray_mask'].float()
    rgb = lambda p: (((p - out['pixel_colors_gt']) ** 2).sum(-1) * m).sum(1) / (m.sum(1) + 1e-3) * 0.01
    sdf_gt = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (40, 40, 40)).astype(np.float32)).to(dev)
    loss = rgb(out['pixel_colors_nr']).sum() + rgb(out['pixel_colors_nr_fine']).sum() + 0.1 * out['sdf_gradient_error'].mean() \
        + F.smooth_l1_loss(out['volume'][0, 0], sdf_gt)
    loss.backward()
    named = dict(net.named_parameters())
    for k in ('nr_net.agg_net.agg_impl.rgb_fc.0.weight', 'nr_net.fine_agg_net.agg_impl.base_fc.0.weight', 'nr_net.image_encoder.conv1.weight',
              'nr_net.fine_dist_decoder.mean_decoder.0.weight', 'nr_net.agg_net.agg_impl.ray_attention.w_qs.weight',
              'nr_net.agg_net.deviation_network.variance'):
        assert named[k].grad is not None and torch.isfinite(named[k].grad).all() and named[k].grad.abs().sum() > 0, k
    c0 = out['pixel_colors_nr'].detach().clone()
    opt.step()
    with torch.no_grad():
        c1 = net(dict(data, eval=True))['pixel_colors_nr']
    assert torch.isfinite(c1).all() and (c1 - c0).abs().max() > 0


def test_trainstep_batched_encoders_equal_scene_by_scene():
    """train.TrainStep with the encoders batched over the scenes of a step (encoder_chunk > 1: one encoder forward / backward
    for the group, InstanceNorm is per image so this is exact) must give the same gradients as the scene-by-scene order."""
    from graspnerf_b200.train import TrainStep
    dev = torch.device('cuda:0')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    batch = []
    for s in range(3):
        scene = make_scene(seed=50 + s, num_views=4, h=96, w=160, radius=0.45)
        ref = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in scene.items() if k not in ('img_feats', 'ray_feats')}
        ref['sdf_gt'] = torch.from_numpy(np.random.default_rng(s).uniform(-1, 1, (40, 40, 40)).astype(np.float32)).to(dev)
        q = {k: torch.from_numpy(v).to(dev) for k, v in make_query(scene, 24, 7 + s).items() if isinstance(v, np.ndarray)}
        batch.append({'step': 0, 'full_vol': True, 'ref_imgs_info': ref, 'que_imgs_info': q, 'src_imgs_info': ref})
    grads = {}
    for chunk in (1, 3):
        net = _mirror_with_golden_weights(dev).train()
        step = TrainStep(net, lr=0.0, encoder_chunk=chunk)
        torch.manual_seed(5)                                  # sample_fine_depth / depth-loss pixels draw from torch's RNG
        step(batch)
        grads[chunk] = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
    assert set(grads[1]) == set(grads[3])
    gmax = max(float(v.double().norm()) for v in grads[1].values())
    worst = 0.0
    for k in grads[1]:
        a, b = grads[3][k].double(), grads[1][k].double()
        if float(b.norm()) < 1e-5 * gmax:
            # e.g. a convolution bias in front of an InstanceNorm: its true gradient is exactly zero, what is left is round-off
            assert float(a.norm()) < 1e-4 * gmax, k
            continue
        rel = float((a - b).norm() / b.norm())
        worst = max(worst, rel)
        assert rel < 1e-2, (k, rel)          # fp32 atomics in gn_k1_backward accumulate in a different order every run (measured worst 3e-3)
    print(f'batched vs scene-by-scene gradients: worst rel-L2 {worst:.2e}')


def _small_train_batch(dev, seeds, with_rays):
    batch = []
    for s in seeds:
        scene = make_scene(seed=50 + s, num_views=4, h=96, w=160, radius=0.45)
        ref = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in scene.items() if k not in ('img_feats', 'ray_feats')}
        ref['sdf_gt'] = torch.from_numpy(np.random.default_rng(s).uniform(-1, 1, (40, 40, 40)).astype(np.float32)).to(dev)
        q = {k: torch.from_numpy(v).to(dev) for k, v in make_query(scene, 24, 7 + s).items() if isinstance(v, np.ndarray)}
        batch.append({'step': 0, 'full_vol': True, 'ref_imgs_info': ref, 'que_imgs_info': q, 'src_imgs_info': ref})
    return batch


@pytest.mark.parametrize('render_rgb', [False, True])
def test_trainstep_cuda_graph_mode(render_rgb):
    """TrainStep(graph=True): forward + losses + backward of a group of scenes replayed as ONE CUDA graph on inputs staged into
    static buffers.  render_rgb off (no random draws in the forward): the graphed step must reproduce the eager step on NEW
    inputs - losses, and at frozen weights the whole gradient bucket - to the noise of the reverse kernels' atomics.  render_rgb on (shipped
    configuration: random fine-sampling offsets and depth-loss pixels come from the graph-safe generator, so the draws differ
    from the eager run): same losses to the sampling noise."""
    import copy
    from graspnerf_b200.train import TrainStep
    dev = torch.device('cuda:0')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    A, B = _small_train_batch(dev, (0, 1), render_rgb), _small_train_batch(dev, (2, 3), render_rgb)
    net_e = _mirror_with_golden_weights(dev).train()
    net_e.nr_net.cfg['render_rgb'] = render_rgb
    if not render_rgb:
        net_e.nr_net.cfg['use_depth_loss'] = False
        for agg in (net_e.nr_net.agg_net, net_e.nr_net.fine_agg_net):          # no RGB head: nothing flips the variance's flag
            agg.deviation_network.fix_s = -1
    net_g = copy.deepcopy(net_e)
    eager, graphed = TrainStep(net_e, lr=1e-3, encoder_chunk=2), TrainStep(net_g, lr=1e-3, encoder_chunk=2, graph=True)
    losses = {}
    for name, step in (('eager', eager), ('graph', graphed)):
        torch.manual_seed(3)
        losses[name] = [step(A), step(B), step(A), step(B)]
    assert graphed.graph_error is None and graphed._g is not None, graphed.graph_error        # the graph was captured and used
    assert all(np.isfinite(losses['graph']))
    tol = 2e-2 if render_rgb else 1e-4
    for le, lg in zip(losses['eager'], losses['graph']):
        assert abs(le - lg) <= tol * max(abs(le), 1e-3), (losses, render_rgb)
    if not render_rgb:
        # gradients of a REPLAY on inputs the graph was not captured with, against the eager step, at identical weights (lr 0)
        ne, ng = copy.deepcopy(net_e), copy.deepcopy(net_e)
        se, sg = TrainStep(ne, lr=0.0, encoder_chunk=2), TrainStep(ng, lr=0.0, encoder_chunk=2, graph=True)
        for step in (se, sg):
            step(A); step(B); step(A)                                   # graph mode: eager, capture on B, replay on A
        assert sg._g is not None
        a, b = sg.bucket.flat[:-1].double(), se.bucket.flat[:-1].double()
        rel = float((a - b).norm() / b.norm())
        assert rel < 2e-3, rel                                          # fp32 atomics of the reverse kernels (order differs run to run)
