"""Torch-tensor front end of the C ABI (include/graspnerf_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every computation below is one of the
hand-written sm_100a kernels.  Nothing in this module has a CPU or eager-PyTorch fallback - without a CUDA device and the
built library these functions raise.
"""
import ctypes as C
from ctypes import byref as C_byref
import numpy as np
import torch

from . import _lib
from .weights import DevicePacker, positional_table, voxel_axis_table

REC_STRIDE, PT_STRIDE, POOL_STRIDE, TOK_STRIDE = 72, 2, 68, 20
# record columns (csrc/gn_common.cuh): ray_feats | dir_diff | rgb, depth | img_feats
REC_RAY, REC_DD, REC_RGB, REC_DEPTH, REC_IMG = slice(0, 32), slice(32, 36), slice(36, 39), 39, slice(40, 72)


def _stream(dev=None):
    """Current stream of the device that owns the tensors (NOT of torch's current device: a scene on cuda:1 must launch on
    cuda:1 even if the caller never called torch.cuda.set_device)."""
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _on(dev):
    """Context: the launchers' per-device caches (kernel attributes, SM count) and the launch itself use cudaGetDevice()."""
    return torch.cuda.device(dev)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f32c(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t, dtype=np.float32))
    return t.to(device=device, dtype=torch.float32).contiguous()


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f'{what} must live on a CUDA device: graspnerf_b200 has no CPU path')


_PACKERS = {}


class HeadWeights:
    """Device-resident packed weights of one (agg_net, dist_decoder) pair + kernel constants.

    `refresh(sd)` re-packs when any source tensor changed (training); inference packs once.  Packing runs on the device
    (weights.DevicePacker: one gather + four tiny fp64 matmuls, no host synchronisation)."""

    def __init__(self, sd, agg_prefix='agg_net.', dd_prefix='dist_decoder.', device='cuda'):
        self.agg_prefix, self.dd_prefix, self.device = agg_prefix, dd_prefix, torch.device(device)
        self._versions = None
        self._pos = {}
        self._axis = {}
        self._var = None
        self.refresh(sd)

    def _keys(self, sd):
        return [k for k in sd if k.startswith(self.agg_prefix) or k.startswith(self.dd_prefix)]

    def refresh(self, sd):
        vers = tuple((k, sd[k]._version, sd[k].data_ptr()) for k in self._keys(sd))
        if vers == self._versions:
            return
        key = (self.agg_prefix, self.dd_prefix, str(self.device))
        if key not in _PACKERS:
            _PACKERS[key] = DevicePacker(sd, self.agg_prefix, self.dd_prefix, self.device)
        self.packer = _PACKERS[key]
        with torch.no_grad():
            self.blob = self.packer.pack({k: sd[k].to(self.device) for k in self.packer.keys})
        # tensor-core operand images (fp16 hi/lo, K-major) + small constants, built on the device from the fp32 blob
        lib = _lib.load()
        self.tc_const = torch.empty(lib.gn_k2a_tc_const_bytes(), dtype=torch.uint8, device=self.device)
        with _on(self.device):
            _lib.check(lib.gn_k2a_tc_prepare(_ptr(self.blob), _ptr(self.tc_const), _stream(self.device)), 'gn_k2a_tc_prepare')
        self._var = sd.get(self.agg_prefix + 'deviation_network.variance')
        if not hasattr(self, 'status'):
            self.status = torch.zeros(1, dtype=torch.int32, device=self.device)     # sticky numerics flag written by K2a
        self._versions = vers

    def check_numerics(self, reset=True):
        """Raises FloatingPointError if any K2a launch with these weights saw a non-finite hit probability, pooled feature or
        token since the last reset: the fp16 hi/lo operand split of the tensor-core kernel overflows for activations >= 65504
        (the fp32 reference would not).  Reading the flag synchronises; call it when convenient (tests, every N training steps)."""
        bad = int(self.status.item())
        if reset:
            self.status.zero_()
        if bad:
            raise FloatingPointError('gn_k2a_forward_tc: non-finite activations (fp16 operand range 65504 exceeded); '
                                     'the weights / features are outside the range the tensor-core path supports')

    @property
    def variance(self):
        """NeuS inverse-std parameter as a python float (inference compositing only: reading it synchronises)."""
        return float(self._var.detach()) if self._var is not None else 0.3

    def pos_table(self, dn):
        if dn not in self._pos:
            self._pos[dn] = torch.from_numpy(positional_table(dn)).to(self.device)
        return self._pos[dn]

    def axis(self, resolution, volume_size=0.3):
        key = (resolution, volume_size)
        if key not in self._axis:
            self._axis[key] = torch.from_numpy(voxel_axis_table(resolution, volume_size)).to(self.device)
        return self._axis[key]


_HW_CACHE = {}


def _cached_head_weights(sd, agg_prefix, dd_prefix, device):
    """HeadWeights for the autograd nodes: one object per (prefix pair, device), re-packed only when a parameter's version
    changed - i.e. once per optimizer step, not once per scene and pass (packing + gn_k2a_tc_prepare are ~10 launches)."""
    key = (agg_prefix, dd_prefix, str(device))
    hw = _HW_CACHE.get(key)
    if hw is None:
        hw = _HW_CACHE[key] = HeadWeights(sd, agg_prefix, dd_prefix, device)
    else:
        hw.refresh(sd)
    return hw


def invalidate_weight_caches():
    """Forget which parameter versions the cached HeadWeights were packed from: the next use re-packs.  train.TrainStep calls
    this right before it captures a CUDA graph, so that the packing kernels are PART of the graph (a replay must pack the
    weights the optimizer has just updated, not find the blob of the capture-time weights)."""
    for hw in _HW_CACHE.values():
        hw._versions = None


def camera_matrices(poses, Ks):
    """[B,V,3,4] K@[R|t] and [B,V,3] camera centres -R^T t; 6 tiny matmuls, done with torch exactly as the reference
    does (render_ops.py:94,112) so that the fp32 values entering the kernel are the reference's."""
    KRt = Ks @ poses
    cam = (-poses[..., :3].transpose(-1, -2) @ poses[..., 3:])[..., 0]
    return KRt.contiguous(), cam.contiguous()


def to_channels_last_maps(fmap):
    """[B,V,C,fh,fw] (any strides) -> contiguous [B,V,fh,fw,C]; free when the encoder already produced channels_last."""
    return fmap.permute(0, 1, 3, 4, 2).contiguous()


class Scene:
    """Batched device-side inputs of the hot path (B scenes x V views).

    Both feature maps are kept in ONE channels-last buffer `feats` [B,V,fh,fw,64] (ray_feats 32 | img_feats 32 per texel),
    so a bilinear tap of K1 is one address and two adjacent 128-byte lines; `ray_feats` / `img_feats` are views of it."""

    def __init__(self, imgs, img_feats, ray_feats, poses, Ks, depth_range, feats_channels_last=False, feats_fused=None):
        # accept single-scene [V,...] tensors like the reference's ref_imgs_info
        if imgs.dim() == 4:
            imgs, poses, Ks, depth_range = imgs[None], poses[None], Ks[None], depth_range[None]
            if feats_fused is not None:
                feats_fused = feats_fused[None]
            else:
                img_feats, ray_feats = img_feats[None], ray_feats[None]
        _require_cuda(imgs, 'imgs')
        dev = imgs.device
        self.device = dev
        self.img_u8 = imgs.dtype == torch.uint8
        if self.img_u8:
            # the planner's images are PNG bytes (main.py:166-171): [B,V,H,W,3] (HWC, as cv2 / imread give them) or already
            # [B,V,H,W,4]; the kernel divides by 255 exactly like color_map_forward, the bytes stay bytes in HBM
            if imgs.shape[-1] == 3:
                imgs = torch.cat([imgs, imgs.new_zeros(imgs.shape[:-1] + (1,))], -1)
            if imgs.shape[-1] != 4:
                raise ValueError('uint8 images must be [B,V,H,W,3] or [B,V,H,W,4]')
            self.imgs = imgs.contiguous()
            self.B, self.V, self.H, self.W, _ = self.imgs.shape
        else:
            imgs = _f32c(imgs, dev)                                                 # [B,V,3,H,W] as the reference holds them
            self.B, self.V, _, self.H, self.W = imgs.shape
            # RGBA-interleaved copy [B,V,H,W,4]: one bilinear tap = one 16-byte texel (layout change only, like channels-last)
            self.imgs = torch.cat([imgs.permute(0, 1, 3, 4, 2), imgs.new_zeros(self.B, self.V, self.H, self.W, 1)], -1).contiguous()
        if feats_fused is not None:   # already [B,V,fh,fw,64]
            self.feats = _f32c(feats_fused, dev)
        else:
            if not feats_channels_last:   # logical [B,V,32,fh,fw] -> [B,V,fh,fw,32] views (no copy yet)
                img_feats, ray_feats = img_feats.permute(0, 1, 3, 4, 2), ray_feats.permute(0, 1, 3, 4, 2)
            if img_feats.shape[-1] != 32 or ray_feats.shape != img_feats.shape:
                raise ValueError('feature maps must be [B,V,fh,fw,32]')
            self.feats = fuse_feature_maps(img_feats.to(dev, torch.float32), ray_feats.to(dev, torch.float32))
        if self.feats.dim() != 5 or self.feats.shape[-1] != 64:
            raise ValueError('fused feature buffer must be [B,V,fh,fw,64]')
        _, _, self.fh, self.fw, _ = self.feats.shape
        self.ray_feats, self.img_feats = self.feats[..., :32], self.feats[..., 32:]
        self.KRt, self.cam = camera_matrices(_f32c(poses, dev), _f32c(Ks, dev))
        self.depth_range = _f32c(depth_range, dev)


def fuse_feature_maps(img_feats_cl, ray_feats_cl):
    """Two channels-last maps [...,fh,fw,32] -> the fused buffer [...,fh,fw,64] (ray_feats | img_feats) K1 consumes.  When the
    arguments are channels-last VIEWS of contiguous NCHW tensors on the GPU (what the encoders produce), one tiled-transpose
    launch (gn_k6_fuse_features); otherwise torch.cat."""
    a, b = ray_feats_cl, img_feats_cl
    if a.is_cuda and b.is_cuda and a.dim() >= 4 and a.shape == b.shape and a.dtype == b.dtype == torch.float32:
        an, bn = a.movedim(-1, -3), b.movedim(-1, -3)                 # back to [...,32,fh,fw]
        if an.is_contiguous() and bn.is_contiguous() and a.shape[-1] == 32:
            fh, fw = a.shape[-3], a.shape[-2]
            planes = an.numel() // (32 * fh * fw)
            if planes <= 65535:
                out = torch.empty(a.shape[:-1] + (64,), dtype=torch.float32, device=a.device)
                with _on(a.device):
                    _lib.check(_lib.load().gn_k6_fuse_features(_ptr(an), _ptr(bn), _ptr(out), planes, fh * fw, _stream(a.device)), 'gn_k6_fuse_features')
                return out
    return torch.cat([ray_feats_cl, img_feats_cl], -1).contiguous()


def images_u8_to_float(imgs_u8, rgba_out=None):
    """uint8 images [V,H,W,3|4] on the GPU -> fp32 [V,3,H,W] = u8 / 255 (color_map_forward, main.py:170) in one launch
    (gn_k6_images_u8); rgba_out (optional uint8 [V,H,W,4]) also receives the RGBA texels K1's uint8 mode gathers."""
    V, H, W, C_ = imgs_u8.shape
    assert imgs_u8.is_cuda and imgs_u8.dtype == torch.uint8 and imgs_u8.is_contiguous() and C_ in (3, 4)
    out = torch.empty((V, 3, H, W), dtype=torch.float32, device=imgs_u8.device)
    if rgba_out is not None:
        assert rgba_out.dtype == torch.uint8 and rgba_out.is_contiguous() and rgba_out.numel() == V * H * W * 4
    with _on(imgs_u8.device):
        _lib.check(_lib.load().gn_k6_images_u8(_ptr(imgs_u8), _ptr(out), _ptr(rgba_out), V, H, W, C_, _stream(imgs_u8.device)), 'gn_k6_images_u8')
    return out


def k1_forward(scene, hw, *, resolution=None, bbox_min=None, volume_size=0.3, pts=None, que_dir=None, dn=None,
               debug_idx=False, ev=None, valid_count=None):
    """K1 launch.  Volume mode: resolution + bbox_min [B,3].  Ray mode: pts [B,N,3], que_dir [B,N/dn,3], dn."""
    lib = _lib.load()
    dev = scene.device
    p = _lib.GnK1Params()
    vol = pts is None
    if vol:
        R = int(resolution)
        N, dn_ = R * R * R, R
        bbox_min = _f32c(bbox_min, dev).reshape(scene.B, 3)
        axis = hw.axis(R, volume_size)
        p.axis, p.bbox_min, p.R = _ptr(axis).value, _ptr(bbox_min).value, R
    else:
        pts = _f32c(pts, dev)
        que_dir = _f32c(que_dir, dev)
        N, dn_ = pts.shape[1], int(dn)
        p.pts, p.que_dir, p.R = _ptr(pts).value, _ptr(que_dir).value, 0
    rec = torch.empty((scene.B, N, scene.V, REC_STRIDE), device=dev, dtype=torch.float32)
    pt = torch.empty((scene.B, N, PT_STRIDE), device=dev, dtype=torch.float32)
    dbg = torch.zeros((scene.B, N, scene.V, 2), device=dev, dtype=torch.int32) if debug_idx else None
    p.imgs, p.ray_feats, p.img_feats = _ptr(scene.imgs).value, scene.feats.data_ptr(), scene.feats.data_ptr() + 128
    p.feat_stride = 64
    p.KRt, p.cam = _ptr(scene.KRt).value, _ptr(scene.cam).value
    p.rec, p.pt, p.dbg_feat_idx = _ptr(rec).value, _ptr(pt).value, _ptr(dbg).value
    p.B, p.V, p.H, p.W, p.fh, p.fw = scene.B, scene.V, scene.H, scene.W, scene.fh, scene.fw
    p.N, p.dn, p.volume_mode, p.img_u8 = N, dn_, 1 if vol else 0, 1 if scene.img_u8 else 0
    if valid_count is not None:                    # [B,V] int32, accumulated by the kernel (renderer.py:174-176 diagnostic, no sync)
        assert valid_count.dtype == torch.int32 and valid_count.numel() == scene.B * scene.V
        p.valid_count = _ptr(valid_count).value
    if ev is not None:
        ev[0].record()
    with _on(dev):
        rc = lib.gn_k1_forward(C.byref(p), _stream(dev))
    if ev is not None:
        ev[1].record()
    _lib.check(rc, 'gn_k1_forward')
    return (rec, pt, dbg) if debug_idx else (rec, pt)


K2A_IMPL = 'tc'      # the tcgen05 kernel.  impl='simt' (fp32 CUDA-core restatement of the same math) exists for the GPU tests only


def k2a_forward(rec, pt, hw, depth_range, *, que_dists=None, dn=1, want_colors=False, debug=False, impl=None,
                want_pooled=True, want_tok=False, resolution=None, bbox_min=None, volume_size=0.3, pts=None, ev=None):
    """K2a launch.  impl 'tc' (tcgen05, the product path) or 'simt' (test cross-check).  want_tok (tc only): also run geometry_fc and emit per-point tokens
    [B,N,20] for the attention-only K2b; needs the points: (resolution, bbox_min) in volume mode or pts [B,N,3]."""
    lib = _lib.load()
    impl = 'tc' if (impl or K2A_IMPL) in ('tc', 'tc3') else 'simt'
    B, N, V, _ = rec.shape
    dev = rec.device
    # the three-tile kernel runs geometry_fc in a second phase that re-reads the pooled rows: it always needs the buffer
    pooled = torch.empty((B, N, POOL_STRIDE), device=dev, dtype=torch.float32) if (want_pooled or (impl == 'tc' and want_tok)) else None
    colors = torch.empty((B, N, 4), device=dev, dtype=torch.float32) if want_colors else None
    dbg = torch.zeros((B, N, V, 8), device=dev, dtype=torch.float32) if debug else None
    tok = None
    p = _lib.GnK2aParams()
    p.rec, p.pt, p.weights, p.depth_range = _ptr(rec).value, _ptr(pt).value, _ptr(hw.blob).value, _ptr(depth_range).value
    if que_dists is not None:
        que_dists = _f32c(que_dists, dev)
    p.que_dists, p.pooled, p.colors, p.dbg_rows = _ptr(que_dists).value, _ptr(pooled).value, _ptr(colors).value, _ptr(dbg).value
    p.B, p.N, p.V, p.dn, p.with_rgb = B, N, V, int(dn), 1 if want_colors else 0
    if impl == 'tc':
        p.tc_const = _ptr(hw.tc_const).value
        p.status = _ptr(hw.status).value
        if want_tok:
            tok = torch.empty((B, N, TOK_STRIDE), device=dev, dtype=torch.float32)
            p.tok = _ptr(tok).value
            if pts is None:
                R = int(resolution)
                bbox_min = _f32c(bbox_min, dev).reshape(B, 3)
                axis = hw.axis(R, volume_size)
                p.axis, p.bbox_min, p.R, p.volume_mode = _ptr(axis).value, _ptr(bbox_min).value, R, 1
            else:
                pts = _f32c(pts, dev)
                p.pts, p.R, p.volume_mode = _ptr(pts).value, 0, 0
    elif want_tok or not want_pooled:
        raise ValueError("tokens are produced by the tensor-core K2a only")
    fn = {'tc': lib.gn_k2a_forward_tc, 'simt': lib.gn_k2a_forward}[impl]
    if ev is not None:
        ev[0].record()
    with _on(dev):
        rc = fn(C.byref(p), _stream(dev))
    if ev is not None:
        ev[1].record()
    _lib.check(rc, f'gn_k2a_forward[{impl}]')
    if want_tok:
        return pooled, colors, dbg, tok
    return pooled, colors, dbg


def k2b_forward(pooled, hw, *, dn, resolution=None, bbox_min=None, volume_size=0.3, pts=None, want_grad=False, tok=None, ev=None):
    """K2b launch.  pooled [B,N,68] -> full head (embed, geometry_fc, attention; grad optional);
    tok [B,N,20] (pooled=None) -> attention-only kernel on K2a-TC's tokens."""
    lib = _lib.load()
    src = pooled if pooled is not None else tok
    B, N, _ = src.shape
    dev = src.device
    p = _lib.GnK2bParams()
    vol = resolution is not None
    if vol:
        R = int(resolution)
        bbox_min = _f32c(bbox_min, dev).reshape(B, 3)
        axis = hw.axis(R, volume_size)
        out = torch.empty((B, 1, R, R, R), device=dev, dtype=torch.float32)
        p.axis, p.bbox_min, p.R = _ptr(axis).value, _ptr(bbox_min).value, R
    else:
        out = torch.empty((B, N), device=dev, dtype=torch.float32)
        if pts is not None:
            pts = _f32c(pts, dev)
        p.pts, p.R = _ptr(pts).value, 0
    grad = torch.empty((B, N, 3), device=dev, dtype=torch.float32) if want_grad else None
    pos = hw.pos_table(int(dn))
    p.pooled, p.tok, p.weights, p.pos_table, p.sdf, p.grad = _ptr(pooled).value, _ptr(tok).value, _ptr(hw.blob).value, _ptr(pos).value, _ptr(out).value, _ptr(grad).value
    p.B, p.N, p.dn, p.volume_mode = B, N, int(dn), 1 if vol else 0
    if ev is not None:
        ev[0].record()
    with _on(dev):
        rc = lib.gn_k2b_forward(C.byref(p), _stream(dev))
    if ev is not None:
        ev[1].record()
    _lib.check(rc, 'gn_k2b_forward')
    return out, grad


def k3_composite(sdf, grad, colors, que_dir, depth, inv_s, cos_anneal_ratio=1.0):
    """sdf [B,rn,dn]; grad [B,rn,dn,3]; colors [B,rn,dn,4]; que_dir [B,rn,3]; depth [B,rn,dn]."""
    lib = _lib.load()
    B, rn, dn = sdf.shape
    dev = sdf.device
    alpha = torch.empty((B, rn, dn), device=dev, dtype=torch.float32)
    hit = torch.empty_like(alpha)
    pix = torch.empty((B, rn, 3), device=dev, dtype=torch.float32)
    rdepth = torch.empty((B, rn), device=dev, dtype=torch.float32)
    eik = torch.empty((B, rn), device=dev, dtype=torch.float32)
    p = _lib.GnK3Params()
    p.sdf, p.grad, p.colors, p.que_dir, p.depth = _ptr(sdf).value, _ptr(grad).value, _ptr(colors).value, _ptr(que_dir).value, _ptr(depth).value
    p.inv_s, p.cos_anneal_ratio = float(inv_s), float(cos_anneal_ratio)
    p.alpha, p.hit_prob, p.pixel_colors, p.render_depth, p.eik_partial = _ptr(alpha).value, _ptr(hit).value, _ptr(pix).value, _ptr(rdepth).value, _ptr(eik).value
    p.B, p.rn, p.dn = B, rn, dn
    with _on(dev):
        _lib.check(lib.gn_k3_composite(C.byref(p), _stream(dev)), 'gn_k3_composite')
    return alpha, hit, pix, rdepth, eik


def k3_coarse_depths(depth_range_q, rn, dn):
    """depth_range_q [B,2] -> [B,rn,dn]  (sample_depth, render_ops.py:146-170, deterministic)."""
    lib = _lib.load()
    B = depth_range_q.shape[0]
    out = torch.empty((B, rn, dn), device=depth_range_q.device, dtype=torch.float32)
    with _on(out.device):
        _lib.check(lib.gn_k3_coarse_depths(_ptr(depth_range_q), _ptr(out), B, rn, dn, _stream(out.device)), 'gn_k3_coarse_depths')
    return out


def k3_fine_depths(depth, hit_prob, depth_range_q, u, want_inds=False):
    """-> sorted fine depths [B,rn,fdn] (+ int64 searchsorted indices)."""
    lib = _lib.load()
    B, rn, dn = depth.shape
    fdn = u.shape[-1]
    out = torch.empty((B, rn, fdn), device=depth.device, dtype=torch.float32)
    inds = torch.empty((B, rn, fdn), device=depth.device, dtype=torch.int64) if want_inds else None
    with _on(out.device):
        _lib.check(lib.gn_k3_fine_depths(_ptr(depth), _ptr(hit_prob), _ptr(depth_range_q), _ptr(u), _ptr(out), _ptr(inds),
                                         B, rn, dn, fdn, _stream(out.device)), 'gn_k3_fine_depths')
    return out, inds


def sample_volume(scene, hw, bbox_min, resolution=40, volume_size=0.3, debug=None, impl=None, valid_count=None):
    """NeuralRayRenderer.sample_volume (renderer.py:164-199) for B scenes: K1 -> K2a -> K2b.  Returns [B,1,R,R,R].
    impl 'tc' (default): tcgen05 K2a emits per-point tokens, K2b is attention-only.  'simt': fp32 CUDA-core K2a + full K2b."""
    impl = 'tc' if (impl or K2A_IMPL) in ('tc', 'tc3') else 'simt'
    rec, pt = k1_forward(scene, hw, resolution=resolution, bbox_min=bbox_min, volume_size=volume_size, valid_count=valid_count)
    if impl == 'tc':
        pooled, _, dbg, tok = k2a_forward(rec, pt, hw, scene.depth_range, debug=debug is not None, impl=impl,
                                          want_pooled=debug is not None, want_tok=True, resolution=resolution,
                                          bbox_min=bbox_min, volume_size=volume_size)
        vol, _ = k2b_forward(None, hw, dn=resolution, resolution=resolution, bbox_min=bbox_min, volume_size=volume_size, tok=tok)
    else:
        pooled, _, dbg = k2a_forward(rec, pt, hw, scene.depth_range, debug=debug is not None, impl='simt')
        vol, _ = k2b_forward(pooled, hw, dn=resolution, resolution=resolution, bbox_min=bbox_min, volume_size=volume_size)
    if debug is not None:
        debug.update(rec=rec, pt=pt, pooled=pooled, rows=dbg)
    return vol


class VolumeGraph:
    """CUDA-graph replay of sample_volume (K1 -> K2a -> K2b) for FIXED device buffers: the three launches and their
    workspaces are captured once, every later call is one cudaGraphLaunch (the host-side launch path - ctypes, allocator,
    three launches - costs more than the kernels at 1 scene/step).  Inputs are whatever `scene` / `bbox_min` hold at
    replay time (same buffers, new contents are fine); weights must not be re-packed after capture."""

    def __init__(self, scene, hw, bbox_min, resolution=40, volume_size=0.3, prologue=None):
        self.scene, self.hw, self.bbox_min = scene, hw, bbox_min
        dev = scene.device if scene is not None else bbox_min.device

        def body():
            sc = prologue() if prologue is not None else scene
            return sample_volume(sc, hw, bbox_min, resolution, volume_size)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                      # warm-up outside capture: kernel attributes, allocator pools
                self.out = body()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.out = body()

    def replay(self):
        """Enqueues the captured launches on the current stream; returns the (static) output tensor [B,1,R,R,R]."""
        self.graph.replay()
        return self.out


# ------------------------------------------------------------------------------------------------ depth-mean head
def depth_mean(ray_feats, coords, img_hw, mean_decoder, fine_mean_decoder=None):
    """predict_mean_for_depth_loss (renderer.py:222-266) in one launch (gn_k3_depth_mean): ray_feats [V,32,fh,fw] (any strides),
    coords int64 [V,num,2], img_hw = (H, W) of the images, mean_decoder / fine_mean_decoder: the nn.Sequential
    Linear-ELU-Linear-ELU-Linear-Softplus of MixtureLogisticsDistDecoder.  -> mean [V,num,2] (, mean_fine [V,num,2])."""
    lib = _lib.load()
    dev = ray_feats.device
    if dev.type != 'cuda' or ray_feats.dtype != torch.float32:
        raise ValueError('depth_mean: fp32 CUDA feature maps required (no CPU path)')
    V, C_, fh, fw = ray_feats.shape
    assert C_ == 32
    coords = coords.to(device=dev, dtype=torch.int64).contiguous()
    num = coords.shape[1]
    H, W = int(img_hw[0]), int(img_hw[1])
    p = _lib.GnDepthMeanParams()
    keep = []

    def six(dec, field):
        lins = [dec[0], dec[2], dec[4]]
        assert tuple(lins[0].weight.shape) == (32, 32) and tuple(lins[1].weight.shape) == (32, 32) and tuple(lins[2].weight.shape) == (2, 32)
        for i, t in enumerate(x for l in lins for x in (l.weight, l.bias)):
            t = t.detach()
            if not t.is_contiguous() or t.dtype != torch.float32 or t.device != dev:
                t = t.to(device=dev, dtype=torch.float32).contiguous()
                keep.append(t)
            field[i] = _ptr(t).value
    six(mean_decoder, p.w_coarse)
    mean = torch.empty((V, num, 2), dtype=torch.float32, device=dev)
    mean_fine = None
    if fine_mean_decoder is not None:
        six(fine_mean_decoder, p.w_fine)
        mean_fine = torch.empty((V, num, 2), dtype=torch.float32, device=dev)
        p.mean_fine = _ptr(mean_fine).value
    p.feats, p.coords, p.mean = _ptr(ray_feats).value, _ptr(coords).value, _ptr(mean).value
    p.stride_v, p.stride_c, p.stride_y, p.stride_x = (int(s) for s in ray_feats.stride())
    p.V, p.num, p.H, p.W, p.fh, p.fw = V, num, H, W, fh, fw
    p.align_corners = 1 if (fh, fw) == (H, W) else 0                              # ops.py:25-33
    with _on(dev):
        _lib.check(lib.gn_k3_depth_mean(C.byref(p), _stream(dev)), 'gn_k3_depth_mean')
    return mean, mean_fine


# ------------------------------------------------------------------------------------------------ RGB head
def ray_setup(coords, poses, Ks, depth_range_q, que_depth, want_rays=False):
    """coords2rays + depth2points + depth2inv_dists (render_ops.py:4-52) in one launch (gn_k3_ray_setup): coords [B,rn,2]
    pixel (x,y), poses [B,3,4], Ks [B,3,3], depth_range_q [B,2], que_depth [B,rn,dn] ->
    pts [B,rn*dn,3], que_dir [B,rn,3], inv_dists [B,rn*dn]  (+ centers [B,rn,3], un-normalised dirs [B,rn,3] with want_rays)."""
    lib = _lib.load()
    B, rn, dn = que_depth.shape
    dev = que_depth.device
    coords, poses, Ks, dr, depth = (_f32c(t, dev) for t in (coords, poses, Ks, depth_range_q, que_depth))
    pts = torch.empty((B, rn * dn, 3), device=dev, dtype=torch.float32)
    que_dir = torch.empty((B, rn, 3), device=dev, dtype=torch.float32)
    inv_dists = torch.empty((B, rn * dn), device=dev, dtype=torch.float32)
    p = _lib.GnRaySetupParams()
    p.coords, p.poses, p.Ks, p.depth, p.depth_range = (_ptr(t).value for t in (coords, poses, Ks, depth, dr))
    p.pts, p.que_dir, p.inv_dists = _ptr(pts).value, _ptr(que_dir).value, _ptr(inv_dists).value
    p.B, p.rn, p.dn = B, rn, dn
    centers = dirs = None
    if want_rays:
        centers, dirs = torch.empty((B, rn, 3), device=dev, dtype=torch.float32), torch.empty((B, rn, 3), device=dev, dtype=torch.float32)
        p.centers, p.dirs = _ptr(centers).value, _ptr(dirs).value
    with _on(dev):
        _lib.check(lib.gn_k3_ray_setup(C.byref(p), _stream(dev)), 'gn_k3_ray_setup')
    return (pts, que_dir, inv_dists, centers, dirs) if want_rays else (pts, que_dir, inv_dists)


def inv_s_from(variance):
    """exp(10 * variance) clipped (neus.py:19, aggregate_net.py:107)."""
    import math
    return min(max(math.exp(float(variance) * 10.0), 1e-6), 1e6)


def render_by_depth(scene, hw, que, que_depth, ray_mask_view_num=2, ray_mask_point_num=8):
    """render_by_depth + network_rendering (renderer.py:90-138), eval mode, for B query views.
    que: dict coords [B,rn,2], poses [B,3,4], Ks [B,3,3], depth_range [B,2] (device tensors); que_depth [B,rn,dn]."""
    B, rn, dn = que_depth.shape
    que_depth = que_depth.contiguous()
    pts, que_dir, inv_dists = ray_setup(que['coords'], que['poses'], que['Ks'], que['depth_range'], que_depth)
    rec, pt = k1_forward(scene, hw, pts=pts, que_dir=que_dir, dn=dn)
    pooled, colors, _ = k2a_forward(rec, pt, hw, scene.depth_range, que_dists=inv_dists, dn=dn, want_colors=True)
    sdf, grad = k2b_forward(pooled, hw, dn=dn, pts=pts, want_grad=True)
    inv_s = inv_s_from(hw.variance)
    alpha, hit, pix, rdepth, eik = k3_composite(sdf.reshape(B, rn, dn), grad.reshape(B, rn, dn, 3),
                                                colors.reshape(B, rn, dn, 4), que_dir, que_depth, inv_s)
    nvalid = pt[..., 0].reshape(B, rn, dn)
    out = {'alpha_values': alpha, 'sdf_values': sdf.reshape(B, rn, dn), 'colors_nr': colors.reshape(B, rn, dn, 4)[..., :3],
           'hit_prob_nr': hit, 'pixel_colors_nr': pix, 'render_depth': rdepth,
           'sdf_gradient_error': (eik.sum(1) / (rn * dn)).reshape(B, 1),
           'ray_mask': (nvalid > ray_mask_view_num).sum(-1) > ray_mask_point_num,             # renderer.py:129-132
           'sdf_grad': grad.reshape(B, rn, dn, 3)}
    return out


def render_rays(scene, hw_coarse, hw_fine, que, dn=40, fdn=40, u=None):
    """render_impl + fine_render_impl (renderer.py:140-162), hierarchical sampling on.  u: [B,rn,fdn] uniform randoms
    (training) or None -> stratified midpoints (eval, render_ops.py:197-200)."""
    B, rn = que['coords'].shape[:2]
    depth = k3_coarse_depths(que['depth_range'], rn, dn)
    out = render_by_depth(scene, hw_coarse, que, depth)
    if u is None:
        u = (0.5 / fdn + torch.arange(fdn, device=depth.device, dtype=torch.float32) / fdn).expand(B, rn, fdn).contiguous()
    fdepth, inds = k3_fine_depths(depth, out['hit_prob_nr'], que['depth_range'], u, want_inds=True)
    fine = render_by_depth(scene, hw_fine, que, fdepth)
    out['depth'] = depth
    for k, v in fine.items():
        out[k + '_fine'] = v
    out['depth_fine'], out['fine_inds'] = fdepth, inds
    return out


# ------------------------------------------------------------------------------------------------ training (volume path)
def k2b_backward(pooled, hw, d_sdf, d_weights, *, dn, resolution=None, bbox_min=None, volume_size=0.3, pts=None):
    """Reverse of k2b_forward (full head).  d_sdf: same layout as the forward output.  Accumulates into d_weights
    (blob layout, float64); returns d_pooled [B,N,68]."""
    assert d_weights.dtype == torch.float64
    lib = _lib.load()
    B, N, _ = pooled.shape
    dev = pooled.device
    p = _lib.GnK2bBwdParams()
    vol = resolution is not None
    if vol:
        R = int(resolution)
        bbox_min = _f32c(bbox_min, dev).reshape(B, 3)
        axis = hw.axis(R, volume_size)
        p.axis, p.bbox_min, p.R = _ptr(axis).value, _ptr(bbox_min).value, R
    else:
        pts = _f32c(pts, dev)
        p.pts, p.R = _ptr(pts).value, 0
    d_sdf = _f32c(d_sdf, dev)
    d_pooled = torch.zeros((B, N, POOL_STRIDE), device=dev, dtype=torch.float32)
    pos = hw.pos_table(int(dn))
    p.pooled, p.weights, p.pos_table, p.d_sdf = _ptr(pooled).value, _ptr(hw.blob).value, _ptr(pos).value, _ptr(d_sdf).value
    p.d_pooled, p.d_weights = _ptr(d_pooled).value, _ptr(d_weights).value
    p.B, p.N, p.dn, p.volume_mode = B, N, int(dn), 1 if vol else 0
    with _on(dev):
        _lib.check(lib.gn_k2b_backward(C.byref(p), _stream(dev)), 'gn_k2b_backward')
    return d_pooled


def k2a_backward(rec, pt, hw, depth_range, d_pooled, d_weights, *, que_dists=None, dn=1, d_colors=None):
    """Reverse of k2a_forward.  Accumulates into d_weights (float64); returns d_rec [B,N,V,64] (gradient of the record's
    ray_feats | img_feats entries).  d_colors [B,N,4] (RGB head): also reverses rgb_fc + the softmax colour blend."""
    assert d_weights.dtype == torch.float64
    lib = _lib.load()
    B, N, V, _ = rec.shape
    dev = rec.device
    d_rec = torch.zeros((B, N, V, 64), device=dev, dtype=torch.float32)
    p = _lib.GnK2aBwdParams()
    if que_dists is not None:
        que_dists = _f32c(que_dists, dev)
    p.rec, p.pt, p.weights, p.depth_range = _ptr(rec).value, _ptr(pt).value, _ptr(hw.blob).value, _ptr(depth_range).value
    p.que_dists, p.d_pooled, p.d_rec, p.d_weights = _ptr(que_dists).value, _ptr(d_pooled).value, _ptr(d_rec).value, _ptr(d_weights).value
    if d_colors is not None:
        d_colors = _f32c(d_colors, dev)
        assert d_colors.shape == (B, N, 4)
    p.d_colors = _ptr(d_colors).value
    p.B, p.N, p.V, p.dn = B, N, V, int(dn)
    with _on(dev):
        _lib.check(lib.gn_k2a_backward(C.byref(p), _stream(dev)), 'gn_k2a_backward')
    return d_rec


def k1_backward(scene, hw, d_rec, *, resolution=None, bbox_min=None, volume_size=0.3, pts=None):
    """Reverse of the two feature gathers of k1_forward: d_rec [B,N,V,64] -> (d_img_feats, d_ray_feats), both
    channels-last [B,V,fh,fw,32]."""
    lib = _lib.load()
    dev = scene.device
    B, N, V, _ = d_rec.shape
    p = _lib.GnK1BwdParams()
    vol = pts is None
    if vol:
        R = int(resolution)
        bbox_min = _f32c(bbox_min, dev).reshape(B, 3)
        axis = hw.axis(R, volume_size)
        p.axis, p.bbox_min, p.R = _ptr(axis).value, _ptr(bbox_min).value, R
    else:
        pts = _f32c(pts, dev)
        p.pts, p.R = _ptr(pts).value, 0
    d_img = torch.zeros(tuple(scene.img_feats.shape), device=scene.device, dtype=torch.float32)
    d_ray = torch.zeros(tuple(scene.ray_feats.shape), device=scene.device, dtype=torch.float32)
    p.KRt, p.d_rec, p.d_img_feats, p.d_ray_feats = _ptr(scene.KRt).value, _ptr(d_rec).value, _ptr(d_img).value, _ptr(d_ray).value
    p.B, p.V, p.H, p.W, p.fh, p.fw, p.N, p.volume_mode = B, V, scene.H, scene.W, scene.fh, scene.fw, N, 1 if vol else 0
    with _on(dev):
        _lib.check(lib.gn_k1_backward(C.byref(p), _stream(dev)), 'gn_k1_backward')
    return d_img, d_ray


def sample_volume_backward(scene, hw, bbox_min, rec, pt, pooled, d_vol, resolution=40, volume_size=0.3):
    """d volume [B,1,R,R,R] -> (d_img_feats, d_ray_feats) channels-last [B,V,fh,fw,32] and the weight gradient in blob
    layout, through gn_k2b_backward -> gn_k2a_backward -> gn_k1_backward (fp32, forward recomputed per kernel)."""
    d_w = torch.zeros(hw.blob.shape, dtype=torch.float64, device=hw.blob.device)     # fp64 accumulators
    d_pooled = k2b_backward(pooled, hw, d_vol, d_w, dn=resolution, resolution=resolution, bbox_min=bbox_min, volume_size=volume_size)
    d_rec = k2a_backward(rec, pt, hw, scene.depth_range, d_pooled, d_w)
    d_img, d_ray = k1_backward(scene, hw, d_rec, resolution=resolution, bbox_min=bbox_min, volume_size=volume_size)
    return d_img, d_ray, d_w.float()


class _SampleVolumeFn(torch.autograd.Function):
    """sample_volume with a hand-written backward: the autograd node the mirror's forward uses in training.
    Forward = K1 -> K2a -> K2b (full head, from the pooled features); the record, the per-point mask and the pooled features are kept
    for the reverse kernels.  Inputs that can carry a gradient: img_feats, ray_feats ([V,32,fh,fw] or [B,V,32,fh,fw])
    and the head parameters (passed as *params in the order of `keys`)."""

    @staticmethod
    def forward(ctx, img_feats, ray_feats, static, keys, *params):
        imgs, poses, Ks, depth_range, bbox_min, R, vs, agg_prefix, dd_prefix = static
        sd = {k: p.detach() for k, p in zip(keys, params)}
        hw = _cached_head_weights(sd, agg_prefix, dd_prefix, img_feats.device)
        scene = Scene(imgs, img_feats.detach(), ray_feats.detach(), poses, Ks, depth_range)
        rec, pt = k1_forward(scene, hw, resolution=R, bbox_min=bbox_min, volume_size=vs)
        pooled, _, _ = k2a_forward(rec, pt, hw, scene.depth_range, impl=K2A_IMPL)      # tcgen05 kernel, pooled features kept
        vol, _ = k2b_forward(pooled, hw, dn=R, resolution=R, bbox_min=bbox_min, volume_size=vs)
        ctx.scene, ctx.hw, ctx.saved = scene, hw, (rec, pt, pooled)
        ctx.meta = (bbox_min, R, vs, agg_prefix, dd_prefix, keys, img_feats.dim() == 4, [p.shape for p in params])
        return vol

    @staticmethod
    def backward(ctx, d_vol):
        bbox_min, R, vs, agg_prefix, dd_prefix, keys, single, shapes = ctx.meta
        rec, pt, pooled = ctx.saved
        d_img, d_ray, d_w = sample_volume_backward(ctx.scene, ctx.hw, bbox_min, rec, pt, pooled, d_vol.contiguous(), R, vs)
        d_img = d_img.permute(0, 1, 4, 2, 3)                       # channels-last -> logical [B,V,32,fh,fw]
        d_ray = d_ray.permute(0, 1, 4, 2, 3)
        if single:
            d_img, d_ray = d_img[0], d_ray[0]
        g = ctx.hw.packer.unpack_grad(d_w)
        pg = tuple(g[k].reshape(s) if (k in g and '.rgb_fc.' not in k) else None for k, s in zip(keys, shapes))   # rgb_fc is not on the volume path
        ctx.scene = ctx.hw = ctx.saved = None
        return (d_img, d_ray, None, None) + pg


def sample_volume_autograd(imgs, img_feats, ray_feats, poses, Ks, depth_range, bbox_min, named_params, resolution=40,
                           volume_size=0.3, agg_prefix='agg_net.', dd_prefix='dist_decoder.'):
    """Differentiable sample_volume (first order): named_params = {reference key: nn.Parameter} of the agg_net.* /
    dist_decoder.* tensors.  Returns volume [B,1,R,R,R] attached to the autograd graph of img_feats, ray_feats and params."""
    keys = tuple(k for k in named_params if k.startswith(agg_prefix) or k.startswith(dd_prefix))
    static = (imgs, poses, Ks, depth_range, bbox_min, int(resolution), volume_size, agg_prefix, dd_prefix)
    return _SampleVolumeFn.apply(img_feats, ray_feats, static, keys, *[named_params[k] for k in keys])


# ------------------------------------------------------------------------------------------------ training (RGB head)
class _RayFeaturesFn(torch.autograd.Function):
    """K1 (ray mode) -> K2a with colours, as an autograd node for the RGB head in training: returns the pooled per-point
    features [B,N,68] (mean32 | var32 | mean_v(w) | nvalid ...) and the blended colours [B,N,4].  The per-ray geometry head
    and the compositing that follow are torch ops (network/ray_head.py) because the reference differentiates them TWICE
    (eikonal term, ibrnet.py:497-504); their cotangents d_pooled / d_colors come back here and go through
    gn_k2a_backward (incl. rgb_fc + softmax blend) -> gn_k1_backward."""

    @staticmethod
    def forward(ctx, img_feats, ray_feats, static, keys, *params):
        imgs, poses, Ks, depth_range, pts, que_dir, inv_dists, dn, agg_prefix, dd_prefix = static
        sd = {k: p.detach() for k, p in zip(keys, params)}
        hw = _cached_head_weights(sd, agg_prefix, dd_prefix, img_feats.device)
        scene = Scene(imgs, img_feats.detach(), ray_feats.detach(), poses, Ks, depth_range)
        rec, pt = k1_forward(scene, hw, pts=pts, que_dir=que_dir, dn=dn)
        pooled, colors, _ = k2a_forward(rec, pt, hw, scene.depth_range, que_dists=inv_dists, dn=dn, want_colors=True)
        ctx.scene, ctx.hw, ctx.saved = scene, hw, (rec, pt, pts, inv_dists)
        ctx.meta = (dn, agg_prefix, dd_prefix, keys, img_feats.dim() == 4, [p.shape for p in params])
        nvalid = pt[..., 0].clone()
        ctx.mark_non_differentiable(nvalid)
        return pooled, colors, nvalid

    @staticmethod
    def backward(ctx, d_pooled, d_colors, _d_nvalid):
        dn, agg_prefix, dd_prefix, keys, single, shapes = ctx.meta
        rec, pt, pts, inv_dists = ctx.saved
        scene, hw = ctx.scene, ctx.hw
        d_w = torch.zeros(hw.blob.shape, dtype=torch.float64, device=hw.blob.device)
        d_rec = k2a_backward(rec, pt, hw, scene.depth_range, d_pooled.contiguous(), d_w, que_dists=inv_dists, dn=dn,
                             d_colors=d_colors.contiguous())
        d_img, d_ray = k1_backward(scene, hw, d_rec, pts=pts)
        d_img, d_ray = d_img.permute(0, 1, 4, 2, 3), d_ray.permute(0, 1, 4, 2, 3)
        if single:
            d_img, d_ray = d_img[0], d_ray[0]
        g = hw.packer.unpack_grad(d_w)
        skip = ('.geometry_fc.', '.ray_attention.', '.out_geometry_fc.')          # per-ray head: gradients come from the torch ops
        pg = tuple(g[k].reshape(s) if (k in g and not any(t in k for t in skip)) else None for k, s in zip(keys, shapes))
        ctx.scene = ctx.hw = ctx.saved = None
        return (d_img, d_ray, None, None) + pg


def ray_features_autograd(imgs, img_feats, ray_feats, poses, Ks, depth_range, pts, que_dir, inv_dists, dn, named_params,
                          agg_prefix='agg_net.', dd_prefix='dist_decoder.'):
    """Differentiable K1 -> K2a of the RGB head: pts [B,N,3] (N = rn*dn), que_dir [B,rn,3], inv_dists [B,N].
    Returns (pooled [B,N,68], colors [B,N,4], nvalid [B,N])."""
    keys = tuple(k for k in named_params if k.startswith(agg_prefix) or k.startswith(dd_prefix))
    static = (imgs, poses, Ks, depth_range, pts.detach().contiguous(), que_dir.detach().contiguous(), inv_dists.detach().contiguous(),
              int(dn), agg_prefix, dd_prefix)
    return _RayFeaturesFn.apply(img_feats, ray_feats, static, keys, *[named_params[k] for k in keys])


# ------------------------------------------------------------------------------------------------ grasp post-processing
def grasp_post(tsdf, qual, rot, width, *, gaussian_filter_sigma=1.0, min_width=1.33, max_width=9.33, tsdf_thres_high=0.5,
               tsdf_thres_low=1e-3, threshold=0.90, max_filter_size=4, max_grasps=512):
    """`process` + `select` of the planner (main.py:23-74) on the device (gn_k4_grasp_post), one scene: tsdf / qual / width
    [R,R,R] (any leading singleton dims), rot [4,R,R,R].  Returns (qual_processed [R,R,R], grasps [max_grasps,9] =
    i, j, k, score, rot0..3, width in np.argwhere order, count int32[1]) - all DEVICE tensors, nothing synchronises."""
    lib = _lib.load()
    dev = tsdf.device
    _require_cuda(tsdf, 'tsdf')
    R = tsdf.shape[-1]
    tsdf, qual, width = (_f32c(t, dev).reshape(R, R, R) for t in (tsdf, qual, width))
    rot = _f32c(rot, dev).reshape(4, R, R, R)
    qual_out = torch.empty((R, R, R), device=dev, dtype=torch.float32)
    scratch = torch.empty((3 * R * R * R,), device=dev, dtype=torch.float32)
    grasps = torch.zeros((max_grasps, 9), device=dev, dtype=torch.float32)
    count = torch.zeros((1,), device=dev, dtype=torch.int32)
    p = _lib.GnGraspPostParams()
    p.tsdf, p.qual, p.rot, p.width = _ptr(tsdf).value, _ptr(qual).value, _ptr(rot).value, _ptr(width).value
    p.qual_out, p.scratch, p.grasps, p.count = _ptr(qual_out).value, _ptr(scratch).value, _ptr(grasps).value, _ptr(count).value
    p.sigma, p.min_width, p.max_width = float(gaussian_filter_sigma), float(min_width), float(max_width)
    p.tsdf_thres_high, p.tsdf_thres_low, p.threshold = float(tsdf_thres_high), float(tsdf_thres_low), float(threshold)
    p.R, p.max_filter_size, p.max_grasps = R, int(max_filter_size), int(max_grasps)
    with _on(dev):
        _lib.check(lib.gn_k4_grasp_post(C.byref(p), _stream(dev)), 'gn_k4_grasp_post')
    return qual_out, grasps, count


# ------------------------------------------------------------------------------------------------ VGN head (K5)
class VgnWeights:
    """Prepared weights of the VGN ConvNet for gn_vgn_forward; re-packed when a source tensor changed (like HeadWeights)."""

    def __init__(self, module):
        self.module, self._versions, self.blob = module, None, None

    def refresh(self):
        sd = dict(self.module.named_parameters())
        vers = tuple((k, v._version, v.data_ptr()) for k, v in sd.items())
        if vers != self._versions:
            from .weights import pack_vgn
            with torch.no_grad():
                self.blob = pack_vgn(sd)
            self._versions = vers
        return self.blob

    def workspace(self, R, dev):
        """Scratch of one gn_vgn_forward call.  A fresh allocation per call (the caching allocator / a capturing graph's private
        pool make it cheap): calls on different streams - engine slots that overlap on the GPU - must not share it."""
        return torch.empty(_lib.load().gn_vgn_workspace_floats(R), dtype=torch.float32, device=dev)


def vgn_forward(volume, vw, out=None):
    """gd/networks.py ConvNet.forward on the device kernels: volume [B,1,R,R,R] -> (qual [B,1,R,R,R], rot [B,4,R,R,R],
    width [B,1,R,R,R]) as views of one [B,6,R,R,R] buffer (`out` optional: [B,>=6,R,R,R]-strided destination)."""
    lib = _lib.load()
    _require_cuda(volume, 'volume')
    dev = volume.device
    B, R = volume.shape[0], volume.shape[-1]
    vol = _f32c(volume, dev).reshape(B, R, R, R)
    blob = vw.refresh()
    if out is None:
        out = torch.empty((B, 6, R, R, R), device=dev, dtype=torch.float32)
    p = _lib.GnVgnParams()
    p.volume, p.weights, p.workspace, p.out = _ptr(vol).value, _ptr(blob).value, _ptr(vw.workspace(R, dev)).value, _ptr(out).value
    p.B, p.R, p.out_scene_stride = B, R, out.stride(0)
    with _on(dev):
        _lib.check(lib.gn_vgn_forward(C.byref(p), _stream(dev)), 'gn_vgn_forward')
    return out[:, 0:1], out[:, 1:5], out[:, 5:6]


# ------------------------------------------------------------------------------------------------ fused encoder stages (K6)
_ACT = {None: 0, 'none': 0, 'relu': 1, 'elu': 2}


class SplitK:
    """Partial outputs [S,N,C,H,W] of a split-K convolution (conv2d_tc(..., allow_split=True)); only norm_act_pad consumes it
    (it adds the partials in a fixed order while it reads)."""

    def __init__(self, parts):
        self.parts = parts


def norm_act_pad(x, norm=None, act=None, pad=0, res=None, res_norm=None, x_pad=0, res_pad=0, want_padded=True, want_unpadded=False):
    """gn_k6_norm_act_pad: out = reflect_pad(act(IN(x) [+ res | + IN(res)]), pad).  x [N,C,H+2*x_pad,W+2*x_pad] (NCHW fp32,
    contiguous) or a SplitK; norm / res_norm: nn.InstanceNorm2d modules (affine) or None.  Returns (padded or None, un-padded or None)."""
    lib = _lib.load()
    splits = None
    if isinstance(x, SplitK):
        splits, x = x.parts, x.parts[0]
        assert norm is not None and x_pad == 0
    dev = x.device
    N, C = x.shape[0], x.shape[1]
    H, W = x.shape[2] - 2 * x_pad, x.shape[3] - 2 * x_pad
    assert x.is_contiguous() and x.dtype == torch.float32
    if not want_padded:
        pad = 0
    outp = torch.empty((N, C, H + 2 * pad, W + 2 * pad), device=dev, dtype=torch.float32) if want_padded else None
    outu = torch.empty((N, C, H, W), device=dev, dtype=torch.float32) if want_unpadded else None
    p = _lib.GnNormActPadParams()
    p.x = _ptr(x).value
    if norm is not None:
        p.gamma, p.beta, p.eps = _ptr(norm.weight).value, _ptr(norm.bias).value, float(norm.eps)
    else:
        p.eps = 1e-5
    if res is not None:
        assert res.is_contiguous() and res.shape[2] - 2 * res_pad == H and res.shape[3] - 2 * res_pad == W and res.shape[1] == C
        p.res = _ptr(res).value
        if res_norm is not None:
            p.res_gamma, p.res_beta = _ptr(res_norm.weight).value, _ptr(res_norm.bias).value
    p.out_padded, p.out_unpadded = _ptr(outp).value, _ptr(outu).value
    if splits is not None:
        p.x_splits, p.x_split_stride = splits.shape[0], splits.stride(0)
    p.N, p.C, p.H, p.W, p.pad, p.x_pad, p.res_pad, p.act = N, C, H, W, int(pad), int(x_pad), int(res_pad), _ACT[act]
    with _on(dev):
        _lib.check(lib.gn_k6_norm_act_pad(C_byref(p), _stream(dev)), 'gn_k6_norm_act_pad')
    return outp, outu


def upsample2x_pad(x, pad=0):
    """gn_k6_upsample2x_pad: F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True) + reflection pad."""
    lib = _lib.load()
    N, C, H, W = x.shape
    assert x.is_contiguous() and x.dtype == torch.float32
    out = torch.empty((N, C, 2 * H + 2 * pad, 2 * W + 2 * pad), device=x.device, dtype=torch.float32)
    with _on(x.device):
        _lib.check(lib.gn_k6_upsample2x_pad(_ptr(x), _ptr(out), N * C, H, W, int(pad), _stream(x.device)), 'gn_k6_upsample2x_pad')
    return out


# ------------------------------------------------------------------------------------------------ encoder convolutions on tcgen05 (K7)
class ConvWeights:
    """fp16 hi/lo operand images of one nn.Conv2d for gn_k7_conv_forward + the k -> input-offset table for a given padded
    input size; rebuilt when the weight tensor changed."""

    def __init__(self, conv):
        self.conv, self._version, self.wimg, self._koff = conv, None, None, {}
        co, ci, kh, kw = conv.weight.shape
        self.K = ci * kh * kw
        self.Kpad = (self.K + 31) // 32 * 32
        self.Npad = (co + 15) // 16 * 16

    def images(self):
        w = self.conv.weight
        ver = (w._version, w.data_ptr())
        if ver != self._version:
            with torch.no_grad():
                co = w.shape[0]
                wk = torch.zeros((self.Npad, self.Kpad), dtype=torch.float32, device=w.device)
                wk[:co, :self.K] = w.detach().reshape(co, self.K)
                hi = wk.to(torch.float16)
                lo = (wk - hi.float()).to(torch.float16)
                def img(t):                                      # [Npad,Kpad] -> [Kpad/32][4][Npad][8]: element (n,k) at unit (k/8)*Npad + n
                    return t.reshape(self.Npad, self.Kpad // 32, 4, 8).permute(1, 2, 0, 3)
                self.wimg = torch.stack([img(hi), img(lo)], 1).contiguous()       # [chunks][2][4][Npad][8]
            self._version = ver
        return self.wimg

    def koff(self, Hp, Wp, dev):
        key = (Hp, Wp, str(dev))
        if key not in self._koff:
            _, ci, kh, kw = self.conv.weight.shape
            k = torch.arange(self.Kpad)
            c, r = k // (kh * kw), k % (kh * kw)
            off = c * (Hp * Wp) + (r // kw) * Wp + (r % kw)
            off[k >= self.K] = 0
            self._koff[key] = off.to(torch.int32).to(dev)
        return self._koff[key]


_CONV_CACHE = {}


K7_SPLIT_POLICY = ((64, 16, 4), (160, 16, 2))       # the 18x32 / 128-channel stage: 4 CTAs per tile; the 36x64 stage: 2


def k7_splits(tiles, nchunks):
    """Split-K factor of a K7 launch: K7_SPLIT_POLICY rows are (max tiles, min k chunks, splits), first match wins; the partial
    outputs cost a pass in the K6 launch that sums them, so only layers with few 128-pixel tiles and a long reduction split."""
    for max_tiles, min_chunks, splits in K7_SPLIT_POLICY:
        if tiles <= max_tiles and nchunks >= min_chunks:
            return splits
    return 1


def conv2d_tc(xp, conv, allow_split=False):
    """F.conv2d(xp, conv.weight, conv.bias, conv.stride, padding=0) on tcgen05 (gn_k7_conv_forward): xp [N,Cin,Hp,Wp] fp32
    contiguous, ALREADY padded for this convolution.  fp16 hi/lo operand split, fp32 accumulation (same scheme as K2a).
    allow_split: layers with few 128-pixel tiles and a long K (the 128-channel stage: 27 tiles, K = 1152) may run split-K
    and return a SplitK of partial outputs for norm_act_pad to sum."""
    lib = _lib.load()
    dev = xp.device
    cw = _CONV_CACHE.get(id(conv))
    if cw is None or cw.conv is not conv:
        cw = _CONV_CACHE[id(conv)] = ConvWeights(conv)
    N, Cin, Hp, Wp = xp.shape
    co, ci, kh, kw = conv.weight.shape
    assert ci == Cin and xp.is_contiguous() and xp.dtype == torch.float32 and conv.stride[0] == conv.stride[1]
    st = conv.stride[0]
    Ho, Wo = (Hp - kh) // st + 1, (Wp - kw) // st + 1
    tiles = (N * Ho * Wo + 127) // 128
    nch = cw.Kpad // 32
    S = k7_splits(tiles, nch) if allow_split else 1
    out = torch.empty(((S,) if S > 1 else ()) + (N, co, Ho, Wo), device=dev, dtype=torch.float32)
    p = _lib.GnConvParams()
    p.ksplit = S
    p.in_, p.wimg, p.koff = _ptr(xp).value, _ptr(cw.images()).value, _ptr(cw.koff(Hp, Wp, dev)).value
    p.bias, p.out = _ptr(conv.bias).value if conv.bias is not None else None, _ptr(out).value
    p.Nimg, p.Cin, p.Hp, p.Wp, p.Cout, p.Npad, p.Ho, p.Wo, p.stride, p.Kpad = N, Cin, Hp, Wp, co, cw.Npad, Ho, Wo, st, cw.Kpad
    with _on(dev):
        _lib.check(lib.gn_k7_conv_forward(C_byref(p), _stream(dev)), 'gn_k7_conv_forward')
    return SplitK(out) if S > 1 else out
