"""Mirror of the reference's operator interface for the path, src/nr/network/render_ops.py: same function names, argument
meaning and result layouts, every one backed by the CUDA kernels (no torch arithmetic on the path; torch only reshapes).

  coords2rays / depth2points / depth2inv_dists / depth2dists   render_ops.py:4-52     gn_k3_ray_setup
  project_points_dict                                          render_ops.py:132-144  gn_k1_forward (ray mode)
  sample_depth / sample_fine_depth                             render_ops.py:146-229  gn_k3_coarse_depths / gn_k3_fine_depths

`project_points_dict` returns the reference's dict ([rfn,qn,rn,dn,C] tensors: dir, depth, mask, ray_feats, rgb) plus what the
kernel produced alongside (img_feats - the reference adds it in get_img_feats - and private handles '_rec' / '_pt' / '_scene'
that let the following stages reuse the record instead of re-assembling it).  'pts' (the projected pixel coordinates) is not
materialised by the fused kernel; nothing on the path reads it after the three bilinear taps.
"""
import torch

from .. import ops


def _dev_dict(info):
    return {k: v for k, v in info.items()}


def coords2rays(coords, poses, Ks):
    """render_ops.py:4-25 -> (centers [qn,rn,3], un-normalised directions [qn,rn,3])."""
    qn, rn, _ = coords.shape
    depth = torch.ones((qn, rn, 1), device=coords.device, dtype=torch.float32)
    dr = torch.tensor([[0.5, 2.0]], device=coords.device).expand(qn, 2).contiguous()
    _, _, _, centers, dirs = ops.ray_setup(coords, poses, Ks, dr, depth, want_rays=True)
    return centers, dirs


def depth2points(que_imgs_info, que_depth):
    """render_ops.py:27-39 -> (que_pts [qn,rn,dn,3], que_dir [qn,rn,dn,3])."""
    qn, rn, dn = que_depth.shape
    pts, que_dir, _ = ops.ray_setup(que_imgs_info['coords'], que_imgs_info['poses'], que_imgs_info['Ks'],
                                    que_imgs_info['depth_range'], que_depth.contiguous())
    return pts.reshape(qn, rn, dn, 3), que_dir.unsqueeze(2).expand(qn, rn, dn, 3)


def depth2inv_dists(depth, depth_range, que_imgs_info=None):
    """render_ops.py:46-52 -> spacings in normalised inverse depth [qn,rn,dn] (last one 1e6).  The kernel computes them next
    to the points; called stand-alone it needs nothing but depth and depth_range (identity camera for the unused outputs)."""
    qn, rn, dn = depth.shape
    dev = depth.device
    eye = torch.eye(3, 4, device=dev).expand(qn, 3, 4).contiguous()
    K = torch.eye(3, device=dev).expand(qn, 3, 3).contiguous()
    coords = torch.zeros((qn, rn, 2), device=dev)
    _, _, inv = ops.ray_setup(coords, eye, K, depth_range, depth.contiguous())
    return inv.reshape(qn, rn, dn)


def sample_depth(depth_range, coords, sample_num, random_sample=False):
    """render_ops.py:146-170, deterministic branch (the only one the reference uses: renderer.py:155 passes False)."""
    if random_sample:
        raise NotImplementedError('random_sample=True is never used by the reference (renderer.py:155)')
    qn, rn, _ = coords.shape
    depth = ops.k3_coarse_depths(depth_range, rn, sample_num)
    dists = torch.cat([depth[..., 1:], torch.full_like(depth[..., :1], 1e6)], -1) - depth
    return depth, dists


def sample_fine_depth(depth, hit_prob, depth_range, sample_num, random_sample, inv_mode=True, u=None):
    """render_ops.py:172-229 (inv_mode only).  Returns the UNSORTED fine depths like the reference (the caller sorts,
    renderer.py:148) - the kernel emits them sorted, which is the same multiset.  u: optional explicit uniforms [qn,rn,fdn]."""
    if not inv_mode:
        raise NotImplementedError('inv_mode=False is never used by the reference')
    qn, rn, _ = depth.shape
    if u is None:
        if random_sample:
            u = torch.rand(qn, rn, sample_num, device=depth.device)
        else:
            u = (0.5 / sample_num + torch.arange(sample_num, device=depth.device, dtype=torch.float32) / sample_num).expand(qn, rn, sample_num)
    fd, _ = ops.k3_fine_depths(depth.contiguous(), hit_prob.detach().contiguous(), depth_range, u.contiguous())
    return fd


class PrjDict(dict):
    """The reference's prj_dict (render_ops.py:139) with the kernel's record riding along."""


def project_points_dict(ref_imgs_info, que_pts, hw=None):
    """render_ops.py:132-144 on K1 (ray mode).  ref_imgs_info needs img_feats AND ray_feats (one fused gather samples both);
    que_pts [qn,rn,dn,3] with qn = 1."""
    qn, rn, dn, _ = que_pts.shape
    if qn != 1:
        raise ValueError('the reference renders one query view per call (qn = 1)')
    scene = ops.Scene(ref_imgs_info.get('imgs_u8', ref_imgs_info['imgs']), ref_imgs_info['img_feats'], ref_imgs_info['ray_feats'],
                      ref_imgs_info['poses'], ref_imgs_info['Ks'], ref_imgs_info['depth_range'])
    dev = scene.device
    pts = que_pts.reshape(1, rn * dn, 3).to(dev, torch.float32).contiguous()
    zero_dir = torch.zeros((1, rn, 3), device=dev)             # dir_diff slot then holds (dir - 0, dir . 0) = (dir, 0)
    rec, pt = ops.k1_forward(scene, hw or _NullAxes(dev), pts=pts, que_dir=zero_dir, dn=dn)
    V = scene.V

    def ref_layout(t):                                          # [1,N,V,C] -> [rfn,qn,rn,dn,C]
        return t[0].permute(1, 0, 2).reshape(V, 1, rn, dn, -1)
    bits = pt[0, :, 1].contiguous().view(torch.int32)
    mask = ((bits[:, None] >> torch.arange(V, device=dev, dtype=torch.int32)[None]) & 1).to(torch.float32)      # [N,V]
    d = PrjDict(dir=ref_layout(rec[..., 32:35]), depth=ref_layout(rec[..., ops.REC_DEPTH:ops.REC_DEPTH + 1]), mask=mask.t().reshape(V, 1, rn, dn, 1),
                ray_feats=ref_layout(rec[..., ops.REC_RAY]), rgb=ref_layout(rec[..., ops.REC_RGB]), img_feats=ref_layout(rec[..., ops.REC_IMG]), pts=None)
    d['_rec'], d['_pt'], d['_scene'], d['_que_pts'] = rec, pt, scene, pts
    return d


class _NullAxes:
    """k1_forward only asks its `hw` argument for the voxel axis table in volume mode; ray mode needs nothing."""

    def __init__(self, device):
        self.device = device
