"""Data-parallel training step of the volume path.

The reference trains on one GPU (trainer.py:77-78 raises for multi_gpus).  Here scenes shard over the ranks
(shard.shard_scenes), every rank runs forward + backward on its own scenes - the hot path through the hand-written CUDA
forward / backward kernels, the 2-D encoders and the VGN head through PyTorch/cuDNN autograd - and the gradients meet in ONE
all-reduce (sum) of a flat fp32 bucket per step (NCCL over NVLink on GPUs, gloo in the CPU tests), divided by the number
of scenes of the global batch; then every rank applies the same Adam update (trainer.py:120-123: Adam, lr 1e-4).

Losses: the four of nrvgn_sdf.yaml (`loss: [render, depth, sdf, vgn]`) restated from their formulas - RenderLoss
(loss.py:50-85, coarse + fine), DepthLoss (loss.py:87-144), SDFLoss' smooth-L1 and eikonal terms (loss.py:149-178) and VGNLoss
(loss.py:194-252).  `training_losses` adds whichever terms the forward produced (render_rgb on / off, true_depth present or
not); the reference's own loss classes work on the mirror's output dict unchanged as well.
"""
import torch
import torch.nn.functional as F


class GradBucket:
    """Flat fp32 buffer holding ALL gradients: one collective per step, static layout (parameters that received no gradient
    in a step - e.g. rgb_fc with render_rgb off - contribute zeros, so every rank sends the same bytes).  Every parameter's
    `.grad` is a VIEW of the flat buffer, so autograd accumulates straight into it (no pack / unpack copies) and the
    optimizer reads the reduced values in place."""

    def __init__(self, params):
        # ALL parameters, also those that do not require a gradient yet: deviation_network.variance starts to after the first
        # training forward (neus.py:16-19); it travels as zeros until then and the bucket layout never changes
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        dev = self.params[0].device
        # one extra trailing element carries this rank's scene count through the SAME all-reduce (uneven shards: the global mean
        # divides by the true number of scenes, not by len(local) * world)
        self.flat = torch.zeros(sum(self.sizes) + 1, dtype=torch.float32, device=dev)
        self.views = [v.view_as(p) for v, p in zip(self.flat[:-1].split(self.sizes), self.params)]
        self.attach()

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def attach(self):
        for v, p in zip(self.views, self.params):
            p.grad = v

    def zero(self):
        """Start of a step: zero the bucket and (re-)attach the views (something may have replaced a .grad)."""
        self.flat.zero_()
        self.attach()

    def pack(self):
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)

    def unpack(self):
        self.attach()

    def allreduce(self, dist=None, scale=1.0, local_count=None):
        """sum over ranks (if a process group is given), then scale.  local_count: this rank's number of scenes; when given,
        the scale is 1 / (sum of the counts over the ranks), computed on the device from the reduced trailing element."""
        self.pack()
        if local_count is not None:
            self.flat[-1:].fill_(float(local_count))          # (a kernel argument; `flat[-1] = x` copies from the host and synchronises)
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        if local_count is not None:
            self.flat[:-1].div_(self.flat[-1].clamp_min(1.0))
        elif scale != 1.0:
            self.flat[:-1].mul_(scale)
        self.unpack()


def sdf_loss(volume, sdf_gt, weight=1.0):
    """SDFLoss smooth-L1 term (loss.py:165-175): voxels whose ground truth is -1 are masked on both sides."""
    valid = (sdf_gt != -1.0).to(volume.dtype)
    return F.smooth_l1_loss(sdf_gt * valid, volume[0, 0] * valid) * weight


def vgn_loss(vgn_pred, grasp_info, weight=1e-2):
    """VGNLoss (loss.py:194-205,219-252): BCE on quality + label * (min-over-symmetry quaternion loss + 0.01 * width MSE)."""
    qual, rot, width = vgn_pred
    _, label, rotations, width_gt = grasp_info
    l_qual = F.binary_cross_entropy(qual, label, reduction='none')
    quat = lambda t: 1.0 - torch.abs(torch.sum(rot * t, dim=1))
    l_rot = torch.min(quat(rotations[:, 0]), quat(rotations[:, 1]))
    l_width = 0.01 * F.mse_loss(width, width_gt, reduction='none')
    return (l_qual + label * (l_rot + l_width)).mean() * weight


def render_loss(out, weight=0.01, fine=True):
    """RenderLoss (loss.py:50-85): ray-masked squared colour error of the coarse pass (+ fine pass, use_nr_fine_loss)."""
    gt, m = out['pixel_colors_gt'], out['ray_mask'].float()

    def one(pr):
        return (torch.sum(((pr - gt) ** 2).sum(-1) * m, 1) / (torch.sum(m, 1) + 1e-3) * weight).sum()
    loss = one(out['pixel_colors_nr'])
    if fine and 'pixel_colors_nr_fine' in out:
        loss = loss + one(out['pixel_colors_nr_fine'])
    return loss


def eikonal_loss(out, weight=0.1):
    """SDFLoss eikonal term (loss.py:172-173): mean over the ray chunks of mean((|d sdf / d pts| - 1)^2), coarse pass."""
    return out['sdf_gradient_error'].mean() * weight


def depth_loss(out, ref, weight=1.0):
    """DepthLoss (loss.py:87-144, l2, synthetic scenes): predicted mean inverse depth of the dist decoder at depth_coords
    vs the ground-truth depth map, both in normalised inverse depth."""
    coords = out['depth_coords'].float()
    depth_maps = ref['true_depth']
    _, _, h, w = depth_maps.shape
    grid = torch.stack([coords[..., 0] / (w - 1) * 2 - 1, coords[..., 1] / (h - 1) * 2 - 1], -1).unsqueeze(1)
    gt = F.grid_sample(depth_maps, grid, mode='bilinear', padding_mode='border', align_corners=True)[:, 0, 0]
    dr = ref['depth_range']
    near, far = -1 / dr[:, 0:1], -1 / dr[:, 1:2]
    gt = torch.clamp((-1 / torch.clamp(gt, min=1e-5) - near) / (far - near), min=0, max=1.0)
    loss = ((gt - out['depth_mean']) ** 2).mean()
    if 'depth_mean_fine' in out:
        loss = loss + ((gt - out['depth_mean_fine']) ** 2).mean()
    return loss * weight


def training_losses(out, data):
    """Sum of the shipped configuration's loss terms that the forward produced."""
    loss = volume_losses(out, data)
    if 'pixel_colors_nr' in out and 'pixel_colors_gt' in out:
        loss = loss + render_loss(out) + eikonal_loss(out)
    if 'depth_mean' in out and 'true_depth' in data['ref_imgs_info']:
        loss = loss + depth_loss(out, data['ref_imgs_info'])
    return loss


def volume_losses(out, data):
    loss = sdf_loss(out['volume'], data['ref_imgs_info']['sdf_gt'])
    if 'grasp_info' in data and 'full_vol' not in data:
        loss = loss + vgn_loss(out['vgn_pred'], data['grasp_info'])
    return loss


def _static_like(x, dev):
    """Device-resident copy of a (nested) sample: tensors / arrays / lists of numbers become device tensors that keep their
    address from step to step (CUDA-graph inputs); ints, strings, None stay as they are."""
    if torch.is_tensor(x):
        return x.detach().to(dev).clone()
    if isinstance(x, dict):
        return {k: _static_like(v, dev) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        if len(x) and all(torch.is_tensor(v) or isinstance(v, dict) for v in x):
            return [_static_like(v, dev) for v in x]
        try:
            return torch.as_tensor(x, dtype=torch.float32).to(dev)           # e.g. bbox3d, a list of lists in the dataset
        except (TypeError, ValueError):
            return x
    return x


def _copy_into(static, new):
    if torch.is_tensor(static):
        if not torch.is_tensor(new):
            new = torch.as_tensor(new, dtype=static.dtype)
        static.copy_(new.reshape(static.shape), non_blocking=True)
    elif isinstance(static, dict):
        for k in static:
            _copy_into(static[k], new[k])
    elif isinstance(static, list):
        for s, n in zip(static, new):
            _copy_into(s, n)


class TrainStep:
    """One optimizer step over a global batch of scenes, this rank's share passed in as a list of `data` dicts."""

    def __init__(self, net, lr=1e-4, dist=None, loss_fn=training_losses, encoder_chunk=8, graph=False):
        """encoder_chunk: the 2-D encoders of up to this many scenes run as ONE batched forward / backward (InstanceNorm is per
        image, so batching is exact): the encoders are ~60 % of the ~6 400 launches of a training scene and the step is
        host-launch bound.  1 = scene by scene (the reference's order of operations).
        graph: capture forward + losses + backward of one group of `encoder_chunk` scenes in ONE CUDA graph (inputs staged into
        static device buffers, gradients accumulate into the flat bucket, random draws through the graph-safe generator) and
        replay it for every full group of a step: ~25 000 launches per group become one cudaGraphLaunch.  The step counters the
        reference keeps on the host (cos_anneal_ratio, fix_s: aggregate_net.py:135-137) are frozen in a captured graph, so graph
        mode is refused (eager step instead) while the configuration makes values depend on them (cos_anneal_end_iter != 0;
        before the variance of SingleVarianceNetwork has started to train, neus.py:16-19)."""
        self.net, self.dist, self.loss_fn, self.encoder_chunk = net, dist, loss_fn, max(1, int(encoder_chunk))
        self.bucket = GradBucket(net.parameters())
        self.opt = torch.optim.Adam(self.bucket.params, lr=lr)
        self.world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
        self.graph = bool(graph)
        self.graph_error = None                          # why graph mode was abandoned (None: not abandoned)
        self._g = None                                   # (signature, static group, CUDAGraph)
        self._total = None
        self.steps_done = 0

    def _group_loss(self, group):
        """One encoder pass for the whole group, then scene by scene through the hot path; the caller runs ONE backward."""
        nr = self.net.nr_net
        imgs = torch.cat([d['ref_imgs_info']['imgs'] for d in group], 0)
        img_f, ray_f = nr.encode({'imgs': imgs}, None, True)
        V = group[0]['ref_imgs_info']['imgs'].shape[0]
        loss = 0.0
        for i, data in enumerate(group):
            d2 = dict(data)
            d2['ref_imgs_info'] = dict(data['ref_imgs_info'], img_feats=img_f[i * V:(i + 1) * V], ray_feats=ray_f[i * V:(i + 1) * V])
            loss = loss + self.loss_fn(self.net(d2), d2)
        return loss

    def _graph_allowed(self):
        nr = getattr(self.net, 'nr_net', None)
        if nr is None or self.steps_done < 1:            # the first step runs eagerly (deviation_network.variance starts to
            return False                                  # require a gradient after the first training forward, neus.py:16-19)
        for agg in (getattr(nr, 'agg_net', None), getattr(nr, 'fine_agg_net', None)):
            if agg is None:
                continue
            if agg.cfg.get('cos_anneal_end_iter'):
                self.graph_error = 'cos_anneal_ratio follows a host-side step counter (cos_anneal_end_iter != 0)'
                return False
            dn = agg.deviation_network
            if dn.fix_s != -1 and not dn.variance.requires_grad:
                return False                              # not yet past fix_s (neus.py:16-19): eager for now, graph later
        return True

    @staticmethod
    def _signature(group):
        def sig(x):
            if torch.is_tensor(x):
                return (tuple(x.shape), str(x.dtype))
            if isinstance(x, dict):
                return tuple((k, sig(v)) for k, v in sorted(x.items()))
            if isinstance(x, (list, tuple)):
                return tuple(sig(v) for v in x) if len(x) and (torch.is_tensor(x[0]) or isinstance(x[0], dict)) else ('list', len(x))
            return type(x).__name__
        return tuple(sig(d) for d in group)

    def _run_group_graphed(self, group):
        dev = self.bucket.flat.device
        sg = self._signature(group)
        if self._g is None or self._g[0] != sg:
            static = []
            for d in group:
                sd = {k: _static_like(v, dev) for k, v in d.items() if k != 'src_imgs_info'}
                if 'src_imgs_info' in d:                  # the dataset passes the reference views again under this key
                    sd['src_imgs_info'] = sd['ref_imgs_info'] if d['src_imgs_info'] is d['ref_imgs_info'] else _static_like(d['src_imgs_info'], dev)
                static.append(sd)
            keep = self.bucket.flat.clone()
            keep_total = self._total.clone()
            try:
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):             # warm-up off the capture: allocator pools, cuDNN plans, kernel attributes
                    for _ in range(2):
                        self._group_loss(static).backward()
                torch.cuda.current_stream(dev).wait_stream(side)
                from . import ops
                ops.invalidate_weight_caches()            # the weight packing must be captured too (replays follow optimizer steps)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    loss = self._group_loss(static)
                    loss.backward()
                    self._total += loss.detach()
            finally:
                torch.cuda.synchronize(dev)
                self.bucket.flat.copy_(keep)              # the warm-up passes accumulated into the bucket: restore it
                self._total.copy_(keep_total)
            self._g = (sg, static, g)
        _, static, g = self._g
        for s, d in zip(static, group):
            for k in s:
                if k == 'src_imgs_info' and s[k] is s.get('ref_imgs_info'):
                    continue
                _copy_into(s[k], d[k])
        g.replay()

    def __call__(self, local_batch):
        self.bucket.zero()
        dev = self.bucket.flat.device
        if self._total is None:
            self._total = torch.zeros((), device=dev)
        self._total.zero_()
        total = self._total
        nr = getattr(self.net, 'nr_net', None)
        chunk = self.encoder_chunk if nr is not None else 1
        use_graph = self.graph and self.graph_error is None and dev.type == 'cuda' and self._graph_allowed()
        for c0 in range(0, len(local_batch), chunk):
            group = local_batch[c0:c0 + chunk]
            same = len({tuple(d['ref_imgs_info']['imgs'].shape) for d in group}) == 1 if chunk > 1 else False
            if chunk > 1 and len(group) > 1 and same:
                if use_graph and len(group) == chunk:
                    # (a failed capture is raised, not papered over: it leaves the CUDA generator in capture state)
                    self._run_group_graphed(group)
                    continue
                loss = self._group_loss(group)
                loss.backward()
                total += loss.detach()
            else:
                for data in group:
                    loss = self.loss_fn(self.net(data), data)
                    loss.backward()
                    total += loss.detach()              # no host synchronisation inside the scene loop
        self.bucket.allreduce(self.dist, local_count=len(local_batch))      # mean over the GLOBAL number of scenes (shards may be uneven)
        self.opt.step()
        self.steps_done += 1
        return float(total) / max(len(local_batch), 1)
