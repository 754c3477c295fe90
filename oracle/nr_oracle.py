"""CPU oracle for the GraspNeRF volumetric-TSDF hot path.  TEST INFRASTRUCTURE ONLY.

This is a from-scratch restatement, in plain torch-CPU tensor arithmetic, of the
algorithm in the reference's src/nr/network (each function cites the reference
file:line it follows).  It exists so that the CUDA kernels in
graspnerf_b200/csrc can be checked on a box that has no /root/reference.

Rules:
  * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
    legs may import this module; the product path never does;
  * pinned against the UNMODIFIED reference executed in the authoring container:
    tests/golden/make_golden.py generates tests/golden/*.npz from the real reference
    and tests/test_oracle_golden.py checks this file against them (forward values, first-order gradients of the volume path, training-mode gradients of the RGB head).  The reference
    ships no tests/golden vectors of its own, so "outputs of the
    reference itself run here" is the pin.

Layout convention used here (differs from the reference's [V,qn,rn,dn,C] tensors):
everything per (point, view) is [N, V, C] with n = r*dn + d (ray r, sample d), and
weights are looked up in a flat state_dict by the reference's own key names.

Index/geometry arithmetic uses a FIXED fp32 operation order without FMA so that the
CUDA kernels (which use __fmul_rn/__fadd_rn/__fdiv_rn for the same expressions) can be
compared BIT-EXACTLY on voxel order, masks and bilinear corner indices.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

ALPHA_GROUND_STATE = -15.0  # renderer.py:32


# --------------------------------------------------------------------------- a1
def voxel_axis_table(resolution=40, volume_size=0.3):
    """1-D voxel-centre coordinates, float64 -> float32 exactly as
    utils/field_utils.py:12-25 ((i * VOXEL + HALF) computed in python double, then
    .astype(np.float32))."""
    voxel = volume_size / resolution
    half = voxel / 2
    return np.array([i * voxel + half for i in range(resolution)]).astype(np.float32)


def voxel_centers(resolution=40, volume_size=0.3):
    """[R^3,3] table, row = i*R*R + j*R + k (x-major, z-minor); field_utils.py:17-27."""
    t = voxel_axis_table(resolution, volume_size)
    g = np.stack(np.meshgrid(t, t, t, indexing='ij'), -1)
    return np.ascontiguousarray(g.reshape(-1, 3))


def volume_query_points(bbox_min, resolution=40, volume_size=0.3, dtype=torch.float32):
    """Query points of sample_volume in (ray r = x*R+y, sample d = R-1-z) order;
    renderer.py:166-170 (add bbox3d[0] in fp32, reshape (1,R*R,R,3), flip dim 2)."""
    pts = torch.from_numpy(voxel_centers(resolution, volume_size)) + torch.tensor(bbox_min, dtype=torch.float32)
    pts = pts.reshape(resolution * resolution, resolution, 3).flip(1)
    return pts.to(dtype).contiguous()  # [rn, dn, 3]


# ----------------------------------------------------------------------- a3, a4
def camera_matrices(poses, Ks):
    """H = K @ [R|t] (3x4, fp32 matmul as render_ops.py:94) and camera centres
    c = -R^T t (render_ops.py:112)."""
    KRt = Ks @ poses
    cam = (-poses[:, :, :3].permute(0, 2, 1) @ poses[:, :, 3:])[..., 0]
    return KRt, cam


def project_points(pts, poses, Ks, h, w):
    """render_ops.py:82-130.  pts [N,3] -> dict of [N,V,*] tensors.
    Fixed op order: x = ((h0*px + h1*py) + h2*pz) + h3, no FMA."""
    KRt, cam = camera_matrices(poses, Ks)
    px, py, pz = pts[:, None, 0], pts[:, None, 1], pts[:, None, 2]   # [N,1]

    def row(i):
        hrow = KRt[None, :, i, :]                                     # [1,V,4]
        return ((hrow[..., 0] * px + hrow[..., 1] * py) + hrow[..., 2] * pz) + hrow[..., 3]
    xc, yc, zc = row(0), row(1), row(2)                               # [N,V]
    near_zero = zc.abs() < 1e-4                                       # render_ops.py:101
    depth = torch.where(near_zero, torch.full_like(zc, 1e-3), zc)     # render_ops.py:102
    u, v = xc / depth, yc / depth                                     # render_ops.py:103
    outside = (u < -0.5) | (u >= w - 0.5) | (v < -0.5) | (v >= h - 0.5)  # render_ops.py:126-127
    mask = (~near_zero) & (~outside)                                  # NB: no z>0 test
    dx = px - cam[None, :, 0]
    dy = py - cam[None, :, 1]
    dz = pz - cam[None, :, 2]
    nrm = torch.sqrt((dx * dx + dy * dy) + dz * dz).clamp_min(1e-5)   # render_ops.py:113-114
    dirs = torch.stack([-dx / nrm, -dy / nrm, -dz / nrm], -1)
    return {'uv': torch.stack([u, v], -1), 'depth': depth, 'mask': mask, 'dir': dirs}


# --------------------------------------------------------------------------- a5
def bilinear_coords(u, size_img, size_map, align_corners):
    """Un-normalised sample coordinate along one axis.  ops.py:29-30 normalises by the
    IMAGE size; F.grid_sample then maps [-1,1] to the MAP grid (align_corners=False uses
    ((x+1)*size-1)/2), padding_mode='border' clamps to [0,size-1].  Returns
    (i0 int64, i1 int64, w0, w1) with i1 clamped in-bounds (its weight is 0 when clamped)."""
    xn = u / (size_img - 1) * 2 - 1
    if align_corners:
        ix = (xn + 1) / 2 * (size_map - 1)
    else:
        ix = ((xn + 1) * size_map - 1) / 2
    ix = ix.clamp(0, size_map - 1)
    f0 = torch.floor(ix)
    w1 = ix - f0
    w0 = (f0 + 1) - ix
    i0 = f0.to(torch.int64)
    i1 = torch.clamp(i0 + 1, max=size_map - 1)
    return i0, i1, w0, w1


def sample_map(fmap, uv, h, w):
    """render_ops.py:54-70 + ops.py:14-34 without the mask multiply.
    fmap [V,C,fh,fw]; uv [N,V,2] -> [N,V,C].  align_corners=True iff map size == image size."""
    V, C, fh, fw = fmap.shape
    ac = (fh == h and fw == w)
    x0, x1, wx0, wx1 = bilinear_coords(uv[..., 0], w, fw, ac)
    y0, y1, wy0, wy1 = bilinear_coords(uv[..., 1], h, fh, ac)
    flat = fmap.permute(0, 2, 3, 1).reshape(V, fh * fw, C)            # [V,fh*fw,C]
    vidx = torch.arange(V)[None, :].expand_as(x0)

    def tap(yi, xi):
        return flat[vidx, yi * fw + xi]                               # [N,V,C]
    # grid_sample order: nw, ne, sw, se  (weights (x1-ix)(y1-iy), (ix-x0)(y1-iy), ...)
    out = tap(y0, x0) * (wx0 * wy0)[..., None]
    out = out + tap(y0, x1) * (wx1 * wy0)[..., None]
    out = out + tap(y1, x0) * (wx0 * wy1)[..., None]
    out = out + tap(y1, x1) * (wx1 * wy1)[..., None]
    return out, (x0, y0)


# ----------------------------------------------------------------------- helpers
def _lin(sd, key, x):
    return F.linear(x, sd[key + '.weight'].to(x.dtype), sd[key + '.bias'].to(x.dtype) if (key + '.bias') in sd else None)


def weighted_mean_var(x, wgt):
    """ibrnet.py:112-116: mean = sum_v x*w ; var = sum_v w*(x-mean)^2   (x [N,V,C], w [N,V,1])."""
    mean = torch.sum(x * wgt, dim=1, keepdim=True)
    var = torch.sum(wgt * (x - mean) ** 2, dim=1, keepdim=True)
    return mean, var


# ----------------------------------------------------------------------- a7, a8
def dist_decode(sd, prefix, ray_feats):
    """dist_decoder.py:99-107 with use_vis=False (nrvgn_sdf.yaml:12-15)."""
    def mlp(name):
        x = F.elu(_lin(sd, f'{prefix}{name}.0', ray_feats))
        x = F.elu(_lin(sd, f'{prefix}{name}.2', x))
        return _lin(sd, f'{prefix}{name}.4', x)
    mean = F.softplus(mlp('mean_decoder'))
    var = F.softplus(mlp('var_decoder')) + 0.05       # AddBias(bias_val) dist_decoder.py:78
    aw = torch.sigmoid(mlp('aw_decoder'))
    return mean, var, aw


def normalised_inverse_depth(depth, depth_range_v):
    """dist_decoder.py:17-24 (is_ref): clamp(depth,1e-5); (-1/d - near)/(far - near)."""
    near = (-1 / depth_range_v[:, 0])[None, :]
    far = (-1 / depth_range_v[:, 1])[None, :]
    d = -1 / torch.clamp(depth, min=1e-5)
    return (d - near) / (far - near)


def ray_probabilities(depth, mean, var, aw, depth_range_v, que_dists=None, dn=None):
    """compute_prob (dist_decoder.py:109-142) + get_near_far_points (6-51), is_ref=True.
    depth [N,V]; mean,var [N,V,2]; aw [N,V,1].  que_dists None -> fixed +-0.005 interval
    (volume mode, dist_decoder.py:121-124); else que_dists [rn,dn] -> half intervals
    (dist_decoder.py:34-38, interval broadcast over views)."""
    d = normalised_inverse_depth(depth, depth_range_v)
    if que_dists is None:
        near, far = d - 0.01 / 2, d + 0.01 / 2
    else:
        half = (que_dists / 2)                                        # [rn,dn]
        ext = torch.cat([half[:, 0:1], half], -1)                     # [rn,dn+1]
        near = d - ext[:, :-1].reshape(-1, 1)
        far = d + ext[:, 1:].reshape(-1, 1)
    mix = torch.cat([aw, 1 - aw], -1)
    d0 = (near[..., None] - mean) * var
    d1 = (far[..., None] - mean) * var
    cdf0 = 0.5 + 0.5 * torch.tanh(d0)
    cdf1 = 0.5 + 0.5 * torch.tanh(d1)
    vis = torch.sum((1 - cdf0) * mix, -1)
    hit = torch.sum((cdf1 - cdf0) * mix, -1)
    eps = 1e-5
    alpha = torch.log(hit / (vis - hit + eps) + eps)
    return alpha, vis, hit


# ---------------------------------------------------------------- a12, a13, a15
def positional_table(n_samples, d_hid=16):
    """ibrnet.py:437-445 sinusoid table [n_samples, d_hid] (float64 -> float32)."""
    pos = np.arange(n_samples)[:, None].astype(np.float64)
    j = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000, 2 * (j // 2) / d_hid)
    tab = ang.copy()
    tab[:, 0::2] = np.sin(ang[:, 0::2])
    tab[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(tab).float()


def embed_points(p):
    """neus.py:21-66 with multires=3: [p, sin p, cos p, sin 2p, cos 2p, sin 4p, cos 4p]."""
    out = [p]
    for f in (1.0, 2.0, 4.0):
        out += [torch.sin(p * f), torch.cos(p * f)]
    return torch.cat(out, -1)


def ray_attention(sd, pfx, g, qmask):
    """MultiHeadAttention(4,16,4,4) ibrnet.py:52-102, ScaledDotProductAttention 7-27.
    g [rn,dn,16]; qmask [rn,dn] (1 = query row attends normally, 0 = row filled with -1e9
    -> uniform softmax; the reference applies the mask on the QUERY axis, ibrnet.py:19-20,88-89)."""
    rn, dn, _ = g.shape
    dt = g.dtype

    def proj(name):
        return F.linear(g, sd[f'{pfx}ray_attention.{name}.weight'].to(dt)).view(rn, dn, 4, 4).transpose(1, 2)
    q, k, v = proj('w_qs'), proj('w_ks'), proj('w_vs')
    att = torch.matmul(q / (4 ** 0.5), k.transpose(2, 3))              # [rn,4,dn,dn]
    att = att.masked_fill(qmask[:, None, :, None] == 0, -1e9)
    att = torch.softmax(att, -1)
    o = torch.matmul(att, v).transpose(1, 2).reshape(rn, dn, 16)
    o = F.linear(o, sd[f'{pfx}ray_attention.fc.weight'].to(dt)) + g
    return F.layer_norm(o, (16,), sd[f'{pfx}ray_attention.layer_norm.weight'].to(dt),
                        sd[f'{pfx}ray_attention.layer_norm.bias'].to(dt), eps=1e-6)


def aggregate(sd, agg_prefix, rec, que_pts, rn, dn, que_dir=None, want_grad=False, want_rgb=True, create_graph=False):
    """NeusAggregationNet._get_embedding (aggregate_net.py:35-70) + IBRNetWithNeuRayNeus.forward
    (ibrnet.py:447-513).  rec: dict of [N,V,*] with rgb, img_feats, ray_feats, dir, mask(float),
    hit_prob, vis (the latter two already mask-multiplied, renderer.py:76-77).
    que_pts [rn,dn,3]; que_dir [rn,dn,3] or None -> (0,0,1) (renderer.py:179).
    Returns dict with sdf [rn,dn], colors [rn,dn,3], grad [rn,dn,3] (if want_grad) and
    intermediates.  create_graph=True is the TRAINING form of ibrnet.py:497-504: sdf and grad stay attached to the
    autograd graph (grad with create_graph=True), so losses on them can be differentiated (second order for grad)."""
    P = agg_prefix
    A = agg_prefix + 'agg_impl.'
    N, V = rec['mask'].shape[:2]
    mask = rec['mask'].reshape(N, V, 1)
    dt = mask.dtype
    if que_dir is None:
        qd = torch.tensor([0.0, 0.0, 1.0], dtype=dt).expand(N, 1, 3)
    else:
        qd = que_dir.reshape(N, 1, 3)
    # aggregate_net.py:47-54
    emb_in = torch.cat([rec['ray_feats'], (rec['hit_prob'][..., None] - 0.5) * 2, (rec['vis'][..., None] - 0.5) * 2], -1)
    prob_emb = _lin(sd, P + 'prob_embed.2', torch.relu(_lin(sd, P + 'prob_embed.0', emb_in)))
    # aggregate_net.py:11-17
    dir_diff = torch.cat([rec['dir'] - qd, torch.sum(rec['dir'] * qd, -1, keepdim=True)], -1)
    # ibrnet.py:457-459
    dfeat = F.elu(_lin(sd, A + 'ray_dir_fc.2', F.elu(_lin(sd, A + 'ray_dir_fc.0', dir_diff))))
    rgb_in = rec['rgb']
    f = torch.cat([rec['rgb'], rec['img_feats']], -1) + dfeat
    # ibrnet.py:466-471
    wgt = mask / (torch.sum(mask, dim=1, keepdim=True) + 1e-8)
    nf = _lin(sd, A + 'neuray_fc.2', F.elu(_lin(sd, A + 'neuray_fc.0', prob_emb)))
    w0 = torch.sigmoid(nf) * wgt
    mean0, var0 = weighted_mean_var(f, w0)
    mean1, var1 = weighted_mean_var(f, wgt)
    gfeat = torch.cat([mean0, var0, mean1, var1], -1)
    # ibrnet.py:474-475
    x = torch.cat([gfeat.expand(-1, V, -1), f, prob_emb], -1)
    x = F.elu(_lin(sd, A + 'base_fc.2', F.elu(_lin(sd, A + 'base_fc.0', x))))
    # ibrnet.py:477-482
    xv = F.elu(_lin(sd, A + 'vis_fc.2', F.elu(_lin(sd, A + 'vis_fc.0', x * wgt))))
    x_res, vis = xv[..., :32], xv[..., 32:]
    vis = torch.sigmoid(vis) * mask
    x = x + x_res
    vis2 = torch.sigmoid(_lin(sd, A + 'vis_fc2.2', F.elu(_lin(sd, A + 'vis_fc2.0', x * vis)))) * mask
    w2 = vis2 / (torch.sum(vis2, dim=1, keepdim=True) + 1e-8)
    mean, var = weighted_mean_var(x, w2)
    pooled = torch.cat([mean[:, 0], var[:, 0], w2.mean(dim=1)], -1)    # [N,65]
    nvalid = torch.sum(mask, dim=1)[:, 0]                              # [N]

    # ---- per-ray geometry head, ibrnet.py:485-504
    pos = positional_table(dn).to(dt)
    pts = que_pts.reshape(rn, dn, 3).to(dt).detach().clone().requires_grad_(want_grad)
    with torch.set_grad_enabled(want_grad or torch.is_grad_enabled()):      # outer grad mode = training-style autograd
        g = torch.cat([pooled.reshape(rn, dn, 65), embed_points(pts)], -1)
        g = F.elu(_lin(sd, A + 'geometry_fc.2', F.elu(_lin(sd, A + 'geometry_fc.0', g))))
        g = g + pos[None]
        g = ray_attention(sd, A, g, (nvalid.reshape(rn, dn) > 1).to(dt))
        sdf = _lin(sd, A + 'out_geometry_fc.1', _lin(sd, A + 'out_geometry_fc.0', g)).clip(-1.0, 1.0)[..., 0]
        sdf = sdf.masked_fill(nvalid.reshape(rn, dn) < 1, 1.0)
        grad = None
        if want_grad:
            grad = torch.autograd.grad(sdf, pts, torch.ones_like(sdf), create_graph=create_graph, retain_graph=create_graph)[0]
    out = {'sdf': sdf if (create_graph or (torch.is_grad_enabled() and not want_grad)) else sdf.detach(), 'grad': grad, 'prob_emb': prob_emb, 'dir_diff': dir_diff, 'f': f,
           'mean1': mean1[:, 0], 'var1': var1[:, 0], 'mean0': mean0[:, 0], 'var0': var0[:, 0],
           'x': x, 'vis2': vis2[..., 0], 'pooled': pooled, 'nvalid': nvalid, 'w0': w0[..., 0]}
    if want_rgb:
        # ibrnet.py:507-511
        r = torch.cat([x, vis2, dir_diff], -1)
        r = F.elu(_lin(sd, A + 'rgb_fc.0', r))
        r = F.elu(_lin(sd, A + 'rgb_fc.2', r))
        r = _lin(sd, A + 'rgb_fc.4', r)
        r = r.masked_fill(mask == 0, -1e9)
        bw = torch.softmax(r, dim=1)
        out['colors'] = torch.sum(rgb_in * bw, dim=1).reshape(rn, dn, 3)
    return out


# ------------------------------------------------------------------ a4+a6+a9 glue
def project_and_sample(scene, pts, with_intermediates=False):
    """project_points_dict (render_ops.py:132-144) + get_img_feats (renderer.py:80-88).
    scene: dict of torch tensors imgs/img_feats/ray_feats/poses/Ks.  pts [N,3]."""
    V, _, h, w = scene['imgs'].shape
    prj = project_points(pts, scene['poses'], scene['Ks'], h, w)
    m = prj['mask'].to(pts.dtype)
    rgb, rgb_idx = sample_map(scene['imgs'], prj['uv'], h, w)
    rayf, feat_idx = sample_map(scene['ray_feats'], prj['uv'], h, w)
    imgf, _ = sample_map(scene['img_feats'], prj['uv'], h, w)
    rec = {'rgb': rgb * m[..., None], 'ray_feats': rayf * m[..., None], 'img_feats': imgf * m[..., None],
           'dir': prj['dir'], 'mask': m, 'depth': prj['depth'], 'uv': prj['uv']}
    if with_intermediates:
        rec['rgb_idx'] = rgb_idx
        rec['feat_idx'] = feat_idx
    return rec


def add_ray_probabilities(sd, dd_prefix, rec, depth_range, que_dists=None):
    """predict_proj_ray_prob (renderer.py:62-78)."""
    mean, var, aw = dist_decode(sd, dd_prefix, rec['ray_feats'])
    alpha, vis, hit = ray_probabilities(rec['depth'], mean, var, aw, depth_range, que_dists)
    m = rec['mask']
    rec['alpha'] = alpha * m + (1 - m) * ALPHA_GROUND_STATE
    rec['vis'] = vis * m
    rec['hit_prob'] = hit * m
    rec['dd_mean'], rec['dd_var'], rec['dd_aw'] = mean, var, aw
    return rec


# --------------------------------------------------------------------------- a2
def sample_volume(sd, scene, resolution=40, volume_size=0.3, dtype=torch.float32, with_intermediates=False):
    """NeuralRayRenderer.sample_volume (renderer.py:164-199), volume_type ['sdf'].
    scene values: torch tensors (+ bbox3d python list).  Returns volume [1,1,R,R,R]
    (and the intermediates dict)."""
    R = resolution
    sc = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in scene.items()}
    sdd = {k: v.to(dtype) for k, v in sd.items()}
    que_pts = volume_query_points(scene['bbox3d'][0], R, volume_size, dtype)          # [R*R,R,3]
    rec = project_and_sample(sc, que_pts.reshape(-1, 3), with_intermediates)
    rec = add_ray_probabilities(sdd, 'dist_decoder.', rec, sc['depth_range'], None)
    agg = aggregate(sdd, 'agg_net.', rec, que_pts, R * R, R, None, want_grad=False, want_rgb=False)
    vol = agg['sdf'].reshape(1, 1, R, R, R).flip(-1)                                  # renderer.py:195-198
    if with_intermediates:
        return vol, rec, agg
    return vol


# ------------------------------------------------------------- a17-a21 (RGB head)
def rays_from_coords(coords, pose, K):
    """coords2rays (render_ops.py:4-25) for one query view: coords [rn,2] pixel (x,y);
    returns centre [3], un-normalised directions [rn,3]."""
    rot_t = pose[:, :3].t()
    centre = -(rot_t @ pose[:, 3:])[:, 0]
    hom = torch.cat([coords, torch.ones_like(coords[:, :1])], 1)
    cam = (torch.inverse(K) @ hom.t())                                   # [3,rn]
    world = (rot_t @ cam + centre[:, None]).t()
    return centre, world - centre[None]


def coarse_depths(depth_range_q, rn, dn):
    """sample_depth (render_ops.py:146-170), deterministic branch (renderer.py:155)."""
    near, far = depth_range_q[0], depth_range_q[1]
    interval = (1 / far - 1 / near) / (dn - 1)
    val = torch.arange(1, dn - 1, dtype=torch.float32)
    ticks = torch.cat([torch.zeros(1), interval * val, (1 / far - 1 / near)[None]])
    depth = 1 / (1 / near + ticks)
    return depth[None].repeat(rn, 1)


def depth_to_dists(depth):
    """depth2dists (render_ops.py:41-44): forward differences, last = 1e6."""
    return torch.cat([depth[..., 1:] - depth[..., :-1], torch.full_like(depth[..., :1], 1e6)], -1)


def depth_to_inv_dists(depth, depth_range_q):
    """depth2inv_dists (render_ops.py:46-52)."""
    near, far = -1 / depth_range_q[0], -1 / depth_range_q[1]
    return depth_to_dists((-1 / depth - near) / (far - near))


def neus_alpha(sdf, grad, que_dir, dists, inv_s, cos_anneal_ratio=1.0):
    """_get_alpha_from_sdf (aggregate_net.py:105-123).  sdf,dists [rn,dn]; grad,que_dir [rn,dn,3]."""
    true_cos = (-que_dir * grad).sum(-1)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + F.relu(-true_cos) * cos_anneal_ratio)
    nxt = sdf + iter_cos * dists * 0.5
    prv = sdf - iter_cos * dists * 0.5
    prev_cdf = torch.sigmoid(prv * inv_s)
    next_cdf = torch.sigmoid(nxt * inv_s)
    return ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0.0, 1.0)


def alpha_to_hit_prob(alpha):
    """alpha_values2hit_prob (render_ops.py:72-80)."""
    t = torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], -1)
    return alpha * torch.cumprod(t, -1)[..., :-1]


def fine_depths(depth, hit_prob, depth_range_q, fdn, u=None):
    """sample_fine_depth (render_ops.py:172-229), inv_mode.  u None -> stratified midpoints
    (eval).  Returns (fine_depth [rn,fdn] UNSORTED, inds int64 [rn,fdn])."""
    near, far = -1 / depth_range_q[0], -1 / depth_range_q[1]
    dinv = (-1 / depth - near) / (far - near)
    centre = (dinv[..., 1:] + dinv[..., :-1]) / 2
    centre = torch.cat([dinv[..., 0:1], centre, dinv[..., -1:]], -1)        # [rn,dn+1]
    hp = hit_prob + 1e-5
    # fixed left-to-right fp32 order for the normaliser and the running sum (torch.sum / torch.cumsum leave the order
    # to the backend; the CUDA sampler follows THIS order so its searchsorted table can be compared bit-exactly)
    tot = hp[..., 0]
    for i in range(1, hp.shape[-1]):
        tot = tot + hp[..., i]
    pdf = hp / tot[..., None]
    run = [pdf[..., 0]]
    for i in range(1, pdf.shape[-1]):
        run.append(run[-1] + pdf[..., i])
    cdf = torch.stack(run, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)              # [rn,dn+1]
    if u is None:
        interval = 1 / fdn
        u = (0.5 * interval + torch.arange(fdn) * interval).expand(cdf.shape[0], fdn)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    b0, b1 = torch.gather(centre, -1, below), torch.gather(centre, -1, above)
    denom = c1 - c0
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - c0) / denom
    fd = b0 + t * (b1 - b0)
    fd = -1 / (fd * (far - near) + near)
    return fd, inds


def render_by_depth(sd, scene, que, que_depth, is_fine, ray_mask_view_num=2, ray_mask_point_num=8, train=False):
    """render_by_depth (renderer.py:110-138) + network_rendering (90-108), eval mode.
    que: dict coords [rn,2], pose [3,4], K [3,3], depth_range [2]; que_depth [rn,dn]."""
    rn, dn = que_depth.shape
    dd = 'fine_dist_decoder.' if is_fine else 'dist_decoder.'
    ag = 'fine_agg_net.' if is_fine else 'agg_net.'
    inv_dists = depth_to_inv_dists(que_depth, que['depth_range'])
    centre, dirs = rays_from_coords(que['coords'], que['pose'], que['K'])
    que_pts = centre[None, None] + dirs[:, None] * que_depth[..., None]          # depth2points render_ops.py:27-39
    que_dir = (-dirs / torch.norm(dirs, dim=1, keepdim=True))[:, None].expand(rn, dn, 3)
    rec = project_and_sample(scene, que_pts.reshape(-1, 3))
    rec = add_ray_probabilities(sd, dd, rec, scene['depth_range'], inv_dists)
    agg = aggregate(sd, ag, rec, que_pts, rn, dn, que_dir, want_grad=True, want_rgb=True, create_graph=train)
    dists = depth_to_dists(que_depth)
    inv_s = torch.exp(sd[ag + 'deviation_network.variance'] * 10.0).clip(1e-6, 1e6)   # neus.py:19, aggregate_net.py:107
    alpha = neus_alpha(agg['sdf'], agg['grad'], que_dir, dists, inv_s)
    hit = alpha_to_hit_prob(alpha)
    out = {'alpha_values': alpha, 'sdf_values': agg['sdf'], 'colors_nr': agg['colors'], 'hit_prob_nr': hit,
           'pixel_colors_nr': torch.sum(hit[..., None] * agg['colors'], 1),
           'sdf_gradient_error': torch.mean((torch.linalg.norm(agg['grad'], ord=2, dim=-1) - 1.0) ** 2).reshape(1, 1),
           'render_depth': torch.sum(hit * que_depth, -1), 'sdf_grad': agg['grad'], 'que_pts': que_pts}
    nv = rec['mask'].reshape(rn, dn, -1).to(torch.int32).sum(-1)                  # renderer.py:129-132
    out['ray_mask'] = (nv > ray_mask_view_num).sum(1) > ray_mask_point_num
    return out


def render_rays(sd, scene, que, dn=40, fdn=40, u=None, train=False, fine_depth=None):
    """render_impl + fine_render_impl (renderer.py:140-162), hierarchical sampling on.  train=True keeps the outputs
    attached to the autograd graph (the samplers see the detached hit probabilities, renderer.py:141)."""
    rn = que['coords'].shape[0]
    depth = coarse_depths(que['depth_range'], rn, dn)
    out = render_by_depth(sd, scene, que, depth, False, train=train)
    with torch.no_grad():
        fd, inds = fine_depths(depth, out['hit_prob_nr'].detach(), que['depth_range'], fdn, u)
        fdepth = torch.sort(fd, -1)[0] if fine_depth is None else fine_depth       # fine_depth: externally supplied samples
    fine = render_by_depth(sd, scene, que, fdepth, True, train=train)
    out['depth'] = depth
    for k, v in fine.items():
        out[k + '_fine'] = v
    out['depth_fine'] = fdepth
    out['fine_inds'] = inds
    return out
