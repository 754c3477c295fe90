"""RGB head in TRAINING: the part of the reference's graph that is differentiated twice.

Reference: IBRNetWithNeuRayNeus.forward ibrnet.py:485-504 (embed -> geometry_fc -> + pos_encoding -> 4-head attention ->
LayerNorm -> out_geometry_fc -> clip -> invalid fill, then torch.autograd.grad(sdf, que_pts, create_graph=True)),
NeusAggregationNet._get_alpha_from_sdf / forward aggregate_net.py:105-140 (NeuS alpha, eikonal error),
alpha_values2hit_prob render_ops.py:72-80 and the colour / depth sums of renderer.py:105-106,136.

The per-(point,view) work - projection, the three bilinear taps, dist decoder, prob_embed, base_fc / vis_fc / vis_fc2,
the cross-view poolings, rgb_fc + softmax blend: > 99 % of the FLOPs - runs in the CUDA kernels with hand-written reverse
kernels (ops.ray_features_autograd: gn_k1_forward / gn_k2a_forward_tc, gn_k2a_backward / gn_k1_backward).  What is left
here is the 16-wide per-ray head over rn x dn tokens and the compositing scan: the eikonal loss (loss.py:172-173)
back-propagates THROUGH d sdf / d pts, i.e. needs the second derivative of this head, which torch autograd provides
exactly as in the reference ("second-order autograd").  Inference does not come through here: it
uses gn_k2b_forward (hand-derived first-order gradient) and gn_k3_composite.
"""
import torch
import torch.nn.functional as F

from .. import ops
from ..weights import positional_table

_POS = {}


def _pos_table(dn, device):
    key = (dn, str(device))
    if key not in _POS:
        _POS[key] = torch.from_numpy(positional_table(dn)).to(device)
    return _POS[key]


def embed_points(p):
    """neus.py:21-66 with multires 3: [p, sin p, cos p, sin 2p, cos 2p, sin 4p, cos 4p] (21 values)."""
    out = [p]
    for f in (1.0, 2.0, 4.0):
        out += [torch.sin(p * f), torch.cos(p * f)]
    return torch.cat(out, -1)


def _ray_attention(att, g, qmask):
    """MultiHeadAttention(4, 16, 4, 4), ibrnet.py:52-102; the mask sits on the QUERY axis (ibrnet.py:19-20,88-89)."""
    rn, dn, _ = g.shape

    def proj(lin):
        return F.linear(g, lin.weight).view(rn, dn, 4, 4).transpose(1, 2)
    q, k, v = proj(att.w_qs), proj(att.w_ks), proj(att.w_vs)
    a = torch.matmul(q / (4 ** 0.5), k.transpose(2, 3))
    a = a.masked_fill(qmask[:, None, :, None] == 0, -1e9)
    a = torch.softmax(a, -1)
    o = torch.matmul(a, v).transpose(1, 2).reshape(rn, dn, 16)
    o = F.linear(o, att.fc.weight) + g
    return F.layer_norm(o, (16,), att.layer_norm.weight, att.layer_norm.bias, eps=1e-6)


def sdf_and_gradient(agg_impl, pooled65, pts, nvalid):
    """ibrnet.py:485-504.  pooled65 [rn,dn,65] (attached to the graph), pts [rn,dn,3], nvalid [rn,dn].
    Returns sdf [rn,dn] and d sum(sdf) / d pts [rn,dn,3] with its graph (create_graph=True)."""
    rn, dn, _ = pooled65.shape
    pts = pts.detach().clone().requires_grad_(True)                   # que_pts.requires_grad_(True), ibrnet.py:486
    with torch.enable_grad():
        gf, og = agg_impl.geometry_fc, agg_impl.out_geometry_fc
        g = torch.cat([pooled65, embed_points(pts)], -1)
        g = F.elu(F.linear(F.elu(F.linear(g, gf[0].weight, gf[0].bias)), gf[2].weight, gf[2].bias))
        g = g + _pos_table(dn, g.device)[None]
        g = _ray_attention(agg_impl.ray_attention, g, (nvalid > 1).to(g.dtype))
        sdf = F.linear(F.linear(g, og[0].weight, og[0].bias), og[1].weight, og[1].bias).clip(-1.0, 1.0)[..., 0]
        sdf = sdf.masked_fill(nvalid < 1, 1.0)
        grad = torch.autograd.grad(sdf, pts, torch.ones_like(sdf), create_graph=True, retain_graph=True)[0]
    return sdf, grad


def neus_alpha(sdf, grad, que_dir, dists, inv_s, cos_anneal_ratio=1.0):
    """aggregate_net.py:105-123.  sdf [rn,dn], grad [rn,dn,3], que_dir [rn,1,3] (unit), dists [rn,dn] (last = 1e6)."""
    true_cos = (-que_dir * grad).sum(-1)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + F.relu(-true_cos) * cos_anneal_ratio)
    nxt = sdf + iter_cos * dists * 0.5
    prv = sdf - iter_cos * dists * 0.5
    prev_cdf, next_cdf = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
    return ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0.0, 1.0)


class _CumprodPositive(torch.autograd.Function):
    """torch.cumprod along the last axis for strictly positive inputs.  Same forward; the backward is the formula
    at::cumprod_backward uses when the input has no zero, reversed_cumsum(grad * out) / x - WITHOUT its `(x == 0).any()` test,
    which synchronises with the host on every call (and cannot be captured in a CUDA graph)."""

    @staticmethod
    def forward(ctx, x):
        out = torch.cumprod(x, -1)
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, out = ctx.saved_tensors
        return torch.flip(torch.cumsum(torch.flip(g * out, [-1]), -1), [-1]) / x


def alpha_to_hit_prob(alpha):
    """render_ops.py:72-80.  (alpha is clipped to [0,1], so every factor 1 - alpha + 1e-10 is >= 1e-10 > 0.)"""
    no_hit = torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], -1)
    return alpha * _CumprodPositive.apply(no_hit)[..., :-1]


def render_by_depth_autograd(nr, ref, que, que_depth, is_fine, is_train=True):
    """render_by_depth + network_rendering (renderer.py:90-138) attached to the autograd graph of the feature maps and the
    head parameters.  ref: ref_imgs_info with img_feats / ray_feats; que: coords [1,rn,2], poses [1,3,4], Ks [1,3,3],
    depth_range [1,2]; que_depth [1,rn,dn] (no gradient: the samplers are not differentiated, renderer.py:141)."""
    B, rn, dn = que_depth.shape
    assert B == 1, 'the reference renders one query view per call (qn = 1)'
    agg = nr.fine_agg_net if is_fine else nr.agg_net
    agg_prefix, dd_prefix = ('fine_agg_net.', 'fine_dist_decoder.') if is_fine else ('agg_net.', 'dist_decoder.')
    que_depth = que_depth.contiguous()
    pts, que_dir, inv_dists = ops.ray_setup(que['coords'], que['poses'], que['Ks'], que['depth_range'], que_depth)   # render_ops.py:4-52
    dists = torch.cat([que_depth[..., 1:] - que_depth[..., :-1], torch.full_like(que_depth[..., :1], 1e6)], -1)   # depth2dists 41-44
    named = {k: v for k, v in nr.named_parameters() if k.startswith((agg_prefix, dd_prefix))}
    pooled, colors, nvalid = ops.ray_features_autograd(ref['imgs'], ref['img_feats'], ref['ray_feats'], ref['poses'], ref['Ks'],
                                                       ref['depth_range'], pts, que_dir, inv_dists, dn, named,
                                                       agg_prefix, dd_prefix)
    nvalid = nvalid[0].reshape(rn, dn)
    sdf, grad = sdf_and_gradient(agg.agg_impl, pooled[0, :, :65].reshape(rn, dn, 65), pts[0].reshape(rn, dn, 3), nvalid)
    # NeusAggregationNet.forward bookkeeping (aggregate_net.py:126-137) and SingleVarianceNetwork.forward (neus.py:16-19)
    if agg.cfg['cos_anneal_end_iter'] and is_train:
        agg.cos_anneal_ratio = min(1.0, agg.step / agg.cfg['cos_anneal_end_iter'])
    if is_train:
        agg.step += 1
        agg.deviation_network.set_step(agg.step)
    dev_net = agg.deviation_network
    if dev_net.fix_s != -1 and dev_net.step > dev_net.fix_s:
        dev_net.variance.requires_grad = True
    inv_s = torch.exp(dev_net.variance * 10.0).clip(1e-6, 1e6)
    alpha = neus_alpha(sdf, grad, que_dir[0][:, None], dists[0], inv_s, agg.cos_anneal_ratio)
    hit = alpha_to_hit_prob(alpha)
    col = colors[0, :, :3].reshape(rn, dn, 3)
    out = {'alpha_values': alpha[None], 'sdf_values': sdf[None], 'colors_nr': col[None], 'hit_prob_nr': hit[None],
           'pixel_colors_nr': torch.sum(hit[..., None] * col, 1)[None],
           'render_depth': torch.sum(hit * que_depth[0], -1)[None],
           'sdf_gradient_error': torch.mean((torch.linalg.norm(grad, ord=2, dim=-1) - 1.0) ** 2).reshape(1, 1),
           's': dev_net.variance.reshape(1, 1),
           'ray_mask': ((nvalid > nr.cfg['ray_mask_view_num']).sum(-1) > nr.cfg['ray_mask_point_num'])[None],
           'sdf_grad': grad[None]}
    return out


def render_rays_autograd(nr, ref, que, dn, fdn, u, is_train=True, fine_depth=None):
    """render_impl + fine_render_impl (renderer.py:140-162) for training: coarse pass, hierarchical resampling from the
    DETACHED coarse hit probabilities (gn_k3_fine_depths), fine pass with the fine_* weights."""
    B, rn = que['coords'].shape[:2]
    depth = ops.k3_coarse_depths(que['depth_range'], rn, dn)
    out = render_by_depth_autograd(nr, ref, que, depth, False, is_train)
    if nr.cfg['use_hierarchical_sampling']:
        if u is None:
            u = (0.5 / fdn + torch.arange(fdn, device=depth.device, dtype=torch.float32) / fdn).expand(B, rn, fdn).contiguous()
        if fine_depth is None:
            fdepth, _ = ops.k3_fine_depths(depth, out['hit_prob_nr'].detach().contiguous(), que['depth_range'], u)
        else:                       # externally supplied fine samples [1,rn,fdn] (parity tests)
            fdepth = fine_depth
        fine = render_by_depth_autograd(nr, ref, que, fdepth, True, is_train)
        for k, v in fine.items():
            out[k + '_fine'] = v
        out['depth_fine'] = fdepth
    out['depth'] = depth
    return out
