"""CPU tests of the host-side pieces of the engines (no CUDA needed): the packed camera block of HostScene, RGB transport
accounting, and the split-K policy table of the encoder convolutions."""
import numpy as np
import torch

from graspnerf_b200 import ops
from graspnerf_b200.engine import HostScene, _small_layout


def _host_scene(channels):
    rng = np.random.default_rng(0)
    V = 6
    u8 = rng.integers(0, 256, (V, 8, 16, channels), dtype=np.uint8)
    poses, Ks = rng.standard_normal((V, 3, 4)).astype(np.float32), rng.standard_normal((V, 3, 3)).astype(np.float32)
    dr, bb = np.tile(np.float32([0.2, 0.8]), (V, 1)), np.float32([-0.15, -0.15, -0.05])
    return HostScene(u8, None, None, poses, Ks, dr, bb), (u8, poses, Ks, dr, bb)


def test_host_scene_packs_the_camera_tensors_into_one_block():
    hs, (u8, poses, Ks, dr, bb) = _host_scene(3)
    assert hs.imgs.shape[-1] == 3 and np.array_equal(hs.imgs.numpy(), u8)            # RGB bytes travel as they are
    for got, want in ((hs.poses, poses), (hs.Ks, Ks), (hs.depth_range, dr), (hs.bbox_min, bb)):
        assert got.shape == want.shape and np.array_equal(got.numpy(), want)
        lo, hi = hs.small.data_ptr(), hs.small.data_ptr() + hs.small.numel() * 4
        assert lo <= got.data_ptr() < hi and (got.data_ptr() - lo) % 16 == 0           # a 16-byte aligned view of the block
    offs, total = _small_layout((hs.poses, hs.Ks, hs.depth_range, hs.bbox_min))
    assert offs == [0, 72, 128, 140] and total == hs.small.numel() == 144
    assert len(hs.tensors()) == 2 and hs.nbytes == u8.size + 144 * 4                  # images + ONE camera block cross PCIe
    hs4, _ = _host_scene(4)
    assert hs4.nbytes - hs.nbytes == 6 * 8 * 16                                       # RGBA ships one more byte per pixel


def test_k7_split_policy_table():
    assert ops.k7_splits(27, 36) == 4            # 18x32 maps of the 128-channel stage (K = 1152)
    assert ops.k7_splits(108, 18) == 2           # 36x64 maps (K = 576)
    assert ops.k7_splits(432, 9) == 1            # 72x128 maps: enough tiles, short reduction
    assert ops.k7_splits(27, 9) == 1             # few tiles but a short reduction (1x1 / stride-2 layers)
    old = ops.K7_SPLIT_POLICY
    try:
        ops.K7_SPLIT_POLICY = ()
        assert ops.k7_splits(27, 36) == 1
    finally:
        ops.K7_SPLIT_POLICY = old
