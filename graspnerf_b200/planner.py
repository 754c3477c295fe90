"""The planner's network call and post-processing with everything on the device between image upload and grasp list.

Mirror of the part of GraspNeRFPlanner (src/nr/main.py:87-253) that surrounds the hot path:
  core(images, extrinsics, intrinsics, depth_range, bbox3d)   main.py:211-253   same arguments, same 5-tuple back
  plan(images, extrinsics, intrinsics, ...)                   main.py:188-209   core -> process -> select -> grasp rows
The reference assembles numpy dicts, converts them with imgs_info_to_torch / to_cuda (pageable host memory, one synchronous
copy per tensor), runs the net and copies FOUR volumes back for scipy (main.py:23-84).  Here the images go through one pinned
staging buffer (uint8 when the caller has the PNG bytes - main.py:166-171 reads uint8 and divides by 255 on the CPU; K1
divides in its gather), the whole forward is one CUDA graph (engine.ForwardEngine) and `process` / `select` run in
gn_k4_grasp_post, so only the grasp rows need to come back.

Simulator-side pieces (Blender renders, file names, Grasp / Transform objects, from_voxel_coordinates) stay with the caller:
`plan` returns plain arrays (voxel index, score, quaternion, width) in the reference's np.argwhere order.
"""
import time

import numpy as np
import torch

from .engine import ForwardEngine, HostScene


class GraspPlanner:
    """net: the mirror `GraspNeRF` (graspnerf_b200.network.name2network['grasp_nerf'](cfg)) with `render_rgb` off
    (main.py:150), already on `device` with its checkpoint loaded (main.py:154-157)."""

    tsdf_thres_high, tsdf_thres_low = 0.0, -0.85                     # main.py:92-93
    bbox3d = [[-0.15, -0.15, -0.0503], [0.15, 0.15, 0.2497]]         # main.py:91

    def __init__(self, net, device='cuda', max_grasps=256):
        self.net = net.eval()
        self.net.nr_net.cfg['render_rgb'] = False
        self.device = torch.device(device)
        self.max_grasps = max_grasps
        self._engines = {}

    def _engine(self, hs):
        key = (tuple(hs.imgs.shape),)
        if key not in self._engines:
            post = dict(tsdf_thres_high=self.tsdf_thres_high, tsdf_thres_low=self.tsdf_thres_low)
            # depth_mean=False: the depth-mean head's outputs (renderer.py:288-289) are never read by the planner (main.py:251-253)
            self._engines[key] = ForwardEngine(self.net, hs, slots=2, device=self.device, post_cfg=post, max_grasps=self.max_grasps,
                                               depth_mean=False)
        return self._engines[key]

    @staticmethod
    def _to_u8(images):
        """uint8 [V,H,W,3] passes through; float images [V,3,H,W] in [0,1] (what core() receives in the reference, after
        color_map_forward) are mapped back to the bytes they came from - exact for images that were uint8 / 255."""
        images = np.asarray(images)
        if images.dtype == np.uint8:
            return images if images.shape[-1] in (3, 4) else np.ascontiguousarray(images.transpose(0, 2, 3, 1))
        u8 = np.rint(images.astype(np.float32) * 255.0)
        if not np.array_equal((u8 / 255.0).astype(np.float32), images.astype(np.float32)):
            raise ValueError('float images must be uint8 / 255 (color_map_forward); pass the uint8 images instead')
        return np.ascontiguousarray(u8.astype(np.uint8).transpose(0, 2, 3, 1))

    def _run(self, images, extrinsics, intrinsics, depth_range, bbox3d):
        imgs = self._to_u8(images)
        _, h, w, _ = imgs.shape
        assert h % 32 == 0 and w % 32 == 0                           # main.py:226
        hs = HostScene(imgs, None, None, np.asarray(extrinsics, np.float32)[:, :3, :], np.asarray(intrinsics, np.float32),
                       np.asarray(depth_range, np.float32), np.asarray(bbox3d, np.float32)[0])
        eng = self._engine(hs)
        t0 = time.time()
        i, _ = eng.submit(hs)
        _, (vols, grasps, count) = eng.collect(i)
        return vols, grasps, count, time.time() - t0

    def core(self, images, extrinsics, intrinsics, depth_range=(0.2, 0.8), bbox3d=None, gt_info=None, que_id=0):
        """main.py:211-253: -> (volume, label, rot, width) numpy arrays [1,1,R,R,R] ([1,4,R,R,R] for rot) and the elapsed time."""
        if gt_info:
            raise NotImplementedError('gt_info (VGNLoss evaluation, main.py:248-249) is outside the hot path')
        vols, _, _, t = self._run(images, extrinsics, intrinsics, depth_range, bbox3d if bbox3d is not None else self.bbox3d)
        v = vols.numpy()
        return v[0][None, None].copy(), v[1][None, None].copy(), v[2:6][None].copy(), v[6][None, None].copy(), t

    def plan(self, images, extrinsics, intrinsics, depth_range=None, bbox3d=None):
        """main.py:188-209 without the simulator-side object construction: -> dict(index int64[G,3], score f32[G],
        rot f32[G,4], width f32[G], planning_time).  Rows are in np.argwhere order (main.py:68-73); the caller permutes /
        converts them (main.py:204-207)."""
        nv = len(images)
        if depth_range is None:
            depth_range = np.asarray([[0.2, 0.8]] * nv, np.float32)          # get_depth_range(fixed=True), main.py:181-183
        _, grasps, count, t = self._run(images, extrinsics, intrinsics, depth_range, bbox3d if bbox3d is not None else self.bbox3d)
        n = min(int(count.item()), self.max_grasps)
        g = grasps[:n].numpy()
        return {'index': g[:, :3].astype(np.int64), 'score': g[:, 3].copy(), 'rot': g[:, 4:8].copy(), 'width': g[:, 8].copy(),
                'count': int(count.item()), 'planning_time': t}
