"""Host-buffer front ends: pinned host scenes in, results out, copies overlapped with compute.

Two engines, both the reference-facing call when the inputs live in HOST memory (what GraspNeRFPlanner.core hands to the
network, src/nr/main.py:225-247):

  VolumeEngine   the hot path alone: images + the encoders' feature maps from the host -> K1 -> K2a -> K2b -> TSDF volume.
  ForwardEngine  the whole GraspNeRF.forward of the planner (eval, render_rgb off as main.py:150): uint8 images from the
                 host -> 2-D encoders (K6 / K7) -> K1 -> K2a -> K2b -> depth-mean head -> VGN head (K5) -> (optional) grasp
                 post-processing on the device (K4); only the image bytes cross PCIe on the way in.

`slots` scenes are kept in flight so the PCIe copies of step i+1 run under the kernels of step i (one copy stream + one
compute stream, ordered by CUDA events; no host synchronisation inside the loop except when a slot is recycled).
ForwardEngine gives every slot a compute stream of its own: a scene's encoders are a dependent chain of ~250 launches on
6 small images that leaves most SMs idle, so the graphs of the scenes in flight overlap on the GPU.

Images may be uint8 ([V,H,W,3] as cv2 / imread give them, or [V,H,W,4]): they cross PCIe as bytes (3 per pixel for RGB; the
slot's graph expands them to 4-byte RGBA texels on the device) and K1 divides by 255 in its gather exactly like
color_map_forward (main.py:170); fp32 [V,3,H,W] images are accepted as the reference holds them.

Result lifetime: a result handed back by submit()/collect()/drain() is a pinned buffer owned by the engine.  Each slot
alternates between TWO output buffers, so a returned buffer stays untouched until the SAME slot has been submitted to twice
more, i.e. for at least `slots` further submit() calls.  Copy it if you need it longer.
"""
import torch

from . import ops


def bind_to_gpu_numa(device_index):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off (sysfs: /sys/bus/pci/devices/<bdf>/numa_node and
    /sys/devices/system/node/node<N>/cpulist) BEFORE it allocates pinned host buffers, so the staging memory of every rank is
    local to its GPU's PCIe root: with 8 ranks on a 2-socket host the H2D traffic otherwise funnels through one socket's
    memory controllers and the inter-socket link (round 1: 7 345 volumes/s at 8 GPUs = 0.59 scaling of the end-to-end leg).
    Returns the node number, or None when the topology is not exposed (single node, container without sysfs) - then nothing
    changes."""
    import ctypes
    import os
    try:
        pr = torch.cuda.get_device_properties(device_index)
        if hasattr(pr, 'pci_bus_id'):
            bdf = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        else:                                               # ask the CUDA runtime torch has already loaded
            buf = ctypes.create_string_buffer(32)
            rt = ctypes.CDLL(None)
            if rt.cudaDeviceGetPCIBusId(buf, 32, int(device_index)) != 0:
                return None
            dom, rest = buf.value.decode().lower().split(':', 1)
            bdf = dom[-4:] + ':' + rest
        node = int(open(f'/sys/bus/pci/devices/{bdf}/numa_node').read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def _pin(x, dtype=None):
    t = torch.as_tensor(x)
    if dtype is not None:
        t = t.to(dtype)
    t = t.contiguous()
    return t.pin_memory() if torch.cuda.is_available() else t


def _small_layout(parts):
    """Offsets (in floats, 16-byte aligned) of the small tensors inside their shared block, and the block's size."""
    offs, o = [], 0
    for p in parts:
        offs.append(o)
        o += (p.numel() + 3) // 4 * 4
    return offs, o


class HostScene:
    """Pinned host buffers of one scene.  imgs: uint8 [V,H,W,3|4] or fp32 [V,3,H,W]; feature maps (VolumeEngine only):
    the two channels-last maps [V,fh,fw,32], interleaved per texel into the fused [V,fh,fw,64] buffer K1 consumes."""

    def __init__(self, imgs, img_feats_cl=None, ray_feats_cl=None, poses=None, Ks=None, depth_range=None, bbox_min=None):
        imgs = torch.as_tensor(imgs)
        if imgs.dtype == torch.uint8:
            # RGB bytes cross PCIe as they are (3 bytes per pixel); the slot expands them to the 4-byte RGBA texels K1 gathers
            # (one bilinear tap = one aligned 32-bit load) with one strided copy on the device
            self.imgs = _pin(imgs)
        else:
            self.imgs = _pin(imgs, torch.float32)
        self.feats = None
        if img_feats_cl is not None:
            self.feats = _pin(ops.fuse_feature_maps(torch.as_tensor(img_feats_cl, dtype=torch.float32),
                                                    torch.as_tensor(ray_feats_cl, dtype=torch.float32)))
        # the four small camera tensors live in ONE pinned block (poses | Ks | depth_range | bbox_min: 564 bytes at V = 6) and
        # cross PCIe as one copy: every cudaMemcpyAsync is a DMA job of its own with a fixed cost of a few microseconds
        parts = [torch.as_tensor(x).to(torch.float32).contiguous() for x in (poses, Ks, depth_range, bbox_min)]
        offs, total = _small_layout(parts)
        block = torch.zeros(total, dtype=torch.float32)
        for p, o in zip(parts, offs):
            block[o:o + p.numel()] = p.reshape(-1)
        self.small = _pin(block)
        self.poses, self.Ks, self.depth_range, self.bbox_min = (self.small[o:o + p.numel()].view(p.shape) for p, o in zip(parts, offs))

    def tensors(self):
        return [t for t in (self.imgs, self.feats, self.small) if t is not None]

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors())


class _Slot:
    def __init__(self, hs, out_shapes, device):
        def dev(t):
            return torch.empty(t.shape, dtype=t.dtype, device=device)
        self.imgs_in = dev(hs.imgs)[None]                 # as shipped: uint8 RGB / RGBA or fp32 [V,3,H,W]
        self.rgb = hs.imgs.dtype == torch.uint8 and hs.imgs.shape[-1] == 3
        self.imgs = torch.zeros(self.imgs_in.shape[:-1] + (4,), dtype=torch.uint8, device=device) if self.rgb else self.imgs_in
        self.feats = dev(hs.feats)[None] if hs.feats is not None else None
        self.small = dev(hs.small)                        # poses | Ks | depth_range | bbox_min, one H2D copy
        parts = (hs.poses, hs.Ks, hs.depth_range, hs.bbox_min)
        views = [self.small[o:o + t.numel()].view(t.shape) for t, o in zip(parts, _small_layout(parts)[0])]
        self.poses, self.Ks, self.depth_range = views[0][None], views[1][None], views[2][None]
        self.bbox_min = views[3].reshape(1, 3)
        # two pinned output sets per slot, alternating (see "Result lifetime" in the module docstring)
        self.out_host = [[torch.empty(s, dtype=d).pin_memory() for s, d in out_shapes] for _ in range(2)]
        self.flip = 0
        self.graph = None
        self.static_out = None
        self.ev_in = torch.cuda.Event()
        self.ev_done = torch.cuda.Event()
        self.busy = False
        self.tag = None
        self.last = None

    def expand_rgb(self):
        """uint8 RGB as shipped -> the RGBA texel buffer (alpha stays 0); part of the slot's CUDA graph."""
        if self.rgb:
            self.imgs[..., :3].copy_(self.imgs_in)


class _Engine:
    """Slot ring shared by both engines; subclasses provide `_out_shapes()` and `_compute(slot) -> list of device tensors`."""

    def __init__(self, example, slots, device, concurrent_slots=False):
        """concurrent_slots: every slot computes on a stream of its own, so the scenes in flight overlap on the GPU (pays when
        a scene's work is a long chain of small launches that cannot fill 148 SMs); otherwise one compute stream, scenes in
        submission order."""
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        shapes = self._out_shapes()
        self.slots = [_Slot(example, shapes, self.device) for _ in range(slots)]
        for s in self.slots:
            s.stream = torch.cuda.Stream(self.device) if concurrent_slots else self.compute_stream
            s.side = torch.cuda.Stream(self.device)
        self.next = 0
        self.h2d_bytes = example.nbytes
        self.d2h_bytes = sum(int(torch.tensor(s).prod()) * torch.empty((), dtype=d).element_size() for s, d in shapes)

    def submit(self, hs, tag=None):
        """Queues one scene.  Returns (slot index, finished) where `finished` is the result of the scene that previously
        occupied the slot - (tag, outputs) - or None."""
        i = self.next
        self.next = (self.next + 1) % len(self.slots)
        s = self.slots[i]
        finished = self.collect(i) if s.busy else None
        with torch.cuda.stream(self.copy_stream):
            for dst, src in ((s.imgs_in, hs.imgs), (s.feats, hs.feats)):
                if dst is not None:
                    dst[0].copy_(src, non_blocking=True)
            s.small.copy_(hs.small, non_blocking=True)
            s.ev_in.record(self.copy_stream)
        with torch.cuda.stream(s.stream):
            s.stream.wait_event(s.ev_in)
            outs = self._compute(s)
            s.flip ^= 1
            hosts = s.out_host[s.flip]                   # the buffer set NOT handed out by the previous collect of this slot
            for h, d in zip(hosts, outs):
                h.copy_(d.reshape(h.shape), non_blocking=True)
            s.ev_done.record(s.stream)
        s.busy, s.tag, s.last = True, tag, hosts
        return i, finished

    def collect(self, i):
        """Blocks until slot i finished; returns (tag, outputs).  Outputs: see the engine's docstring."""
        s = self.slots[i]
        s.ev_done.synchronize()
        s.busy = False
        return s.tag, self._wrap(s.last)

    def drain(self):
        return [self.collect(i) for i, s in enumerate(self.slots) if s.busy]

    def _wrap(self, hosts):
        return hosts[0] if len(hosts) == 1 else hosts


class VolumeEngine(_Engine):
    """Hot path with host buffers.  Result per scene: the pinned host volume [1,1,R,R,R]."""

    def __init__(self, head_weights, example, resolution=40, volume_size=0.3, slots=3, device='cuda'):
        if example.feats is None:
            raise ValueError('VolumeEngine needs the feature maps in the HostScene (use ForwardEngine for images-in)')
        self.hw, self.R, self.vs = head_weights, resolution, volume_size
        super().__init__(example, slots, device)

    def _out_shapes(self):
        return [((1, 1, self.R, self.R, self.R), torch.float32)]

    def _compute(self, s):
        if s.graph is None:                   # first use of the slot: capture (layout prep + K1 + K2a + K2b) once
            def prologue(s=s):
                s.expand_rgb()
                return ops.Scene(s.imgs, None, None, s.poses, s.Ks, s.depth_range, feats_fused=s.feats)
            s.graph = ops.VolumeGraph(None, self.hw, s.bbox_min, self.R, self.vs, prologue=prologue)
        return [s.graph.replay()]


class ForwardEngine(_Engine):
    """The planner's whole network call with host buffers (GraspNeRFPlanner.core, main.py:211-253): images in, volumes out.

    net: the mirror `GraspNeRF` (graspnerf_b200.network), on `device`, eval mode.  Per scene the device runs image_encoder /
    init_net / vis_encoder (renderer.py:275-279; K6 / K7), sample_volume (K1 -> K2a -> K2b), the depth-mean head
    (renderer.py:288-289: always on in eval) and the VGN head (renderer.py:323-330), captured in ONE CUDA graph per slot
    when capture succeeds (eager otherwise; `self.graphed` says which).  Result per scene, pinned host tensors:
      [volumes [7,R,R,R] = tsdf, qual, rot0..3, width,  grasps [max_grasps,9],  count int32[1]]
    (grasps / count from gn_k4_grasp_post - main.py:23-74 - when post_cfg is given, else zeros)."""

    def __init__(self, net, example, slots=3, device='cuda', post_cfg=None, max_grasps=256, depth_mean=True, use_graph=True,
                 concurrent_slots=True):
        self.net = net.eval()
        self.R = net.nr_net.cfg['volume_resolution']
        self.post_cfg, self.max_grasps, self.depth_mean, self.use_graph = post_cfg, max_grasps, depth_mean, use_graph
        self.graphed = None
        if example.imgs.dtype != torch.uint8:
            raise ValueError('ForwardEngine takes uint8 images [V,H,W,3|4] (the planner reads PNG bytes, main.py:166-171)')
        super().__init__(example, slots, device, concurrent_slots=concurrent_slots)

    def _out_shapes(self):
        R = self.R
        return [((7, R, R, R), torch.float32), ((self.max_grasps, 9), torch.float32), ((1,), torch.int32)]

    def _body(self, s):
        nr = self.net.nr_net
        # color_map_forward + transpose (main.py:170,192) and the RGBA texels of K1 in one launch
        imgs = ops.images_u8_to_float(s.imgs_in[0], s.imgs if s.rgb else None)
        ref = {'imgs': imgs, 'imgs_u8': s.imgs, 'poses': s.poses[0], 'Ks': s.Ks[0], 'depth_range': s.depth_range[0],
               'bbox3d': s.bbox_min.reshape(1, 3)}
        coords = None
        if self.depth_mean:
            # the depth-mean head's random pixels (a sort-based randperm, ~0.1 ms of small launches) depend on nothing: side stream
            cur = torch.cuda.current_stream(self.device)
            s.side.wait_stream(cur)
            with torch.cuda.stream(s.side):
                coords = nr.draw_depth_coords(imgs.shape[0], imgs.shape[2], imgs.shape[3], imgs.device)
        ref['img_feats'], ref['ray_feats'] = nr.encode(ref, ref, False)
        vol = nr.sample_volume(ref)
        if self.depth_mean:
            cur.wait_stream(s.side)
            s.depth = nr.predict_mean_for_depth_loss(ref, coords=coords)     # renderer.py:288-289 (outputs stay on the device)
        R = self.R
        vols = torch.empty((7, R, R, R), device=vol.device, dtype=torch.float32)
        vols[0].copy_(vol.reshape(R, R, R))
        self.net.vgn_net(vol, out=vols[1:].unsqueeze(0))             # K5 writes qual | rot | width straight into the result buffer
        if self.post_cfg is not None:
            _, grasps, count = ops.grasp_post(vols[0], vols[1], vols[2:6], vols[6], max_grasps=self.max_grasps, **self.post_cfg)
        else:
            grasps = torch.zeros((self.max_grasps, 9), device=vols.device)
            count = torch.zeros((1,), device=vols.device, dtype=torch.int32)
        return [vols, grasps, count]

    def _compute(self, s):
        with torch.no_grad():
            if not self.use_graph or self.graphed is False:
                return self._body(s)
            if s.graph is None:
                cur = torch.cuda.current_stream(self.device)
                try:
                    for _ in range(2):                    # warm-up outside capture: cuDNN algorithm selection, allocator pools
                        self._body(s)
                    cur.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=cur):
                        s.static_out = self._body(s)
                    s.graph, self.graphed = g, True
                except Exception:                          # an op in the torch-side modules refused capture: stay eager
                    torch.cuda.synchronize(self.device)
                    s.graph, self.graphed = None, False
                    return self._body(s)
            s.graph.replay()
            return s.static_out
