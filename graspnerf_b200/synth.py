"""Deterministic synthetic scenes for parity tests and bench.py.

Follows the recipe in SURVEY.md section 8d: V cameras on the simulator's ring
(reference: src/gd/simulation.py:351-357, size=0.3 at simulation.py:28),
intrinsics scaled from src/nr/dataset/database.py:92-109, depth_range [0.2, 0.8],
bbox3d [[-.15,-.15,-.05],[.15,.15,.25]].  All random tensors come from numpy's
PCG64 Generator (stream-stable across numpy versions), never torch's RNG, so the
GPU box regenerates bit-identical inputs without shipping them.
"""
import math
import numpy as np

BBOX3D = [[-0.15, -0.15, -0.05], [0.15, 0.15, 0.25]]
DEPTH_RANGE = (0.2, 0.8)


def intrinsics(h, w):
    """Pinhole K for an h x w image: the 720x1280 camera (fx=fy=892.62,
    cx=639.5, cy=359.5) scaled by h/720 (database.py:92-109 uses 0.4 -> 288x512)."""
    s = h / 720.0
    return np.array([[892.62 * s, 0.0, 639.5 * s],
                     [0.0, 892.62 * s, 359.5 * s],
                     [0.0, 0.0, 1.0]], dtype=np.float32)


def look_at_pose(eye, target, up=(0.0, 0.0, 1.0)):
    """OpenCV world->camera [R|t] (x right, y down, z forward)."""
    eye = np.asarray(eye, np.float64)
    fwd = np.asarray(target, np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, np.float64))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], 0)
    t = -R @ eye
    return np.concatenate([R, t[:, None]], 1).astype(np.float32)


def ring_poses(num_views, radius=0.6, theta=math.pi / 6, center=(0.0, 0.0, -0.05), phase=0.0):
    poses = []
    for i in range(num_views):
        phi = 2.0 * math.pi * i / num_views + phase
        eye = np.asarray(center) + radius * np.array(
            [math.sin(theta) * math.cos(phi), math.sin(theta) * math.sin(phi), math.cos(theta)])
        poses.append(look_at_pose(eye, center))
    return np.stack(poses, 0)


def make_scene(seed=0, num_views=6, h=288, w=512, feat_dim=32, feat_scale=4, phase=None,
               radius=0.6, theta=math.pi / 6):
    """Returns a dict of float32 numpy arrays:
    imgs [V,3,h,w] in [0,1); img_feats, ray_feats [V,C,h/4,w/4] ~ N(0,1);
    poses [V,3,4]; Ks [V,3,3]; depth_range [V,2]; bbox3d (python list)."""
    rng = np.random.default_rng(seed)
    fh, fw = h // feat_scale, w // feat_scale
    if phase is None:
        phase = 0.0 if seed == 0 else float(rng.uniform(0, 2 * math.pi))
    scene = {
        'imgs': rng.random((num_views, 3, h, w), dtype=np.float32),
        'img_feats': rng.standard_normal((num_views, feat_dim, fh, fw), dtype=np.float32),
        'ray_feats': rng.standard_normal((num_views, feat_dim, fh, fw), dtype=np.float32),
        'poses': ring_poses(num_views, radius=radius, theta=theta, phase=phase),
        'Ks': np.repeat(intrinsics(h, w)[None], num_views, 0),
        'depth_range': np.repeat(np.array([DEPTH_RANGE], np.float32), num_views, 0),
        'bbox3d': [list(BBOX3D[0]), list(BBOX3D[1])],
    }
    return scene


def make_query(scene, num_rays=512, seed=0, view=0):
    """Query-view info for the RGB head: coords are pixel (x, y) of `view`."""
    rng = np.random.default_rng(seed + 1000003)
    _, _, h, w = scene['imgs'].shape
    xs = rng.integers(0, w, size=num_rays)
    ys = rng.integers(0, h, size=num_rays)
    coords = np.stack([xs, ys], -1).astype(np.float32)[None]
    return {
        'imgs': scene['imgs'][view:view + 1],
        'poses': scene['poses'][view:view + 1],
        'Ks': scene['Ks'][view:view + 1],
        'coords': coords,
        'depth_range': scene['depth_range'][view:view + 1],
    }
