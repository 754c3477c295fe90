"""Data-parallel training step of the volume path.

The reference trains on one GPU (trainer.py:77-78 raises for multi_gpus).  Here scenes shard over the ranks
(shard.shard_scenes), every rank runs forward + backward on its own scenes - the hot path through the hand-written CUDA
forward / backward kernels, the 2-D encoders and the VGN head through PyTorch/cuDNN autograd - and the gradients meet in ONE
all-reduce (sum) of a flat fp32 bucket per step (NCCL over NVLink on GPUs, gloo in the CPU tests), divided by the number
of scenes of the global batch; then every rank applies the same Adam update (trainer.py:120-123: Adam, lr 1e-4).

Losses: the two of nrvgn_sdf.yaml that depend only on the volume path - SDFLoss' smooth-L1 term (loss.py:165-175) and VGNLoss
(loss.py:194-252) - restated from their formulas.  The render / eikonal / depth terms need the RGB head's backward, which
is not implemented yet (render_rgb must be off).
"""
import torch
import torch.nn.functional as F


class GradBucket:
    """Flat fp32 view of all gradients: one collective per step, static layout (parameters that received no gradient in a
    step - e.g. deviation_network.variance, rgb_fc - contribute zeros, so every rank sends the same bytes)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(self.sizes), dtype=torch.float32, device=dev)
        self.views = [v.view_as(p) for v, p in zip(self.flat.split(self.sizes), self.params)]

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def pack(self):
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)

    def unpack(self):
        for v, p in zip(self.views, self.params):
            p.grad = v          # the optimizer reads the bucket's memory directly

    def allreduce(self, dist=None, scale=1.0):
        """sum over ranks (if a process group is given), then scale (1 / global number of scenes)."""
        self.pack()
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        if scale != 1.0:
            self.flat.mul_(scale)
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


def volume_losses(out, data):
    loss = sdf_loss(out['volume'], data['ref_imgs_info']['sdf_gt'])
    if 'grasp_info' in data and 'full_vol' not in data:
        loss = loss + vgn_loss(out['vgn_pred'], data['grasp_info'])
    return loss


class TrainStep:
    """One optimizer step over a global batch of scenes, this rank's share passed in as a list of `data` dicts."""

    def __init__(self, net, lr=1e-4, dist=None, loss_fn=volume_losses):
        self.net, self.dist, self.loss_fn = net, dist, loss_fn
        self.bucket = GradBucket(net.parameters())
        self.opt = torch.optim.Adam(self.bucket.params, lr=lr)
        self.world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1

    def __call__(self, local_batch):
        self.opt.zero_grad(set_to_none=True)
        total = 0.0
        for data in local_batch:
            loss = self.loss_fn(self.net(data), data)
            loss.backward()
            total += float(loss.detach())
        n_global = len(local_batch) * self.world
        self.bucket.allreduce(self.dist, 1.0 / n_global)
        self.opt.step()
        return total / max(len(local_batch), 1)
