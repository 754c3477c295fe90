"""CPU tests of the host-side pieces of train.TrainStep's CUDA-graph mode and of the sync-free cumprod of the RGB head."""
import numpy as np
import torch

from graspnerf_b200.network.ray_head import _CumprodPositive, alpha_to_hit_prob
from graspnerf_b200.train import TrainStep, _copy_into, _static_like


def test_cumprod_positive_matches_torch_cumprod_forward_and_backward():
    """ray_head._CumprodPositive = torch.cumprod for strictly positive inputs (render_ops.py:72-80 feeds 1 - alpha + 1e-10),
    with the backward torch uses when the input has no zero - minus its host-synchronising zero test."""
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(3, 7, 41, generator=g, dtype=torch.float64) * 0.999 + 1e-10).requires_grad_(True)
    w = torch.randn(3, 7, 41, generator=g, dtype=torch.float64)
    (torch.cumprod(x, -1) * w).sum().backward()
    want, x.grad = x.grad.clone(), None
    out = _CumprodPositive.apply(x)
    assert torch.equal(out, torch.cumprod(x, -1))
    (out * w).sum().backward()
    assert torch.allclose(x.grad, want, rtol=1e-12, atol=1e-14)
    # the formula of the reference through the new node: alpha in [0, 1] incl. the end points
    alpha = torch.tensor([[0.0, 0.3, 1.0, 0.5, 0.0]], dtype=torch.float32, requires_grad=True)
    hit = alpha_to_hit_prob(alpha)
    ref = alpha * torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], -1), -1)[..., :-1]
    assert torch.equal(hit, ref)
    hit.sum().backward()
    assert torch.isfinite(alpha.grad).all()


def test_static_staging_keeps_addresses_and_copies_values():
    dev = torch.device('cpu')
    ref = {'imgs': torch.rand(2, 3, 4, 4), 'bbox3d': [[-0.15, -0.15, -0.05], [0.15, 0.15, 0.25]], 'name': 'scene'}
    data = {'step': 3, 'ref_imgs_info': ref, 'grasp_info': [torch.arange(6).reshape(2, 3), torch.rand(2)]}
    st = _static_like(data, dev)
    assert st['step'] == 3 and st['ref_imgs_info']['name'] == 'scene'
    assert torch.is_tensor(st['ref_imgs_info']['bbox3d']) and st['ref_imgs_info']['bbox3d'].shape == (2, 3)
    assert st['ref_imgs_info']['imgs'].data_ptr() != ref['imgs'].data_ptr()
    ptrs = (st['ref_imgs_info']['imgs'].data_ptr(), st['grasp_info'][0].data_ptr(), st['ref_imgs_info']['bbox3d'].data_ptr())
    new = {'step': 4, 'ref_imgs_info': {'imgs': torch.rand(2, 3, 4, 4), 'bbox3d': [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], 'name': 'other'},
           'grasp_info': [torch.arange(6).reshape(2, 3) + 10, torch.rand(2)]}
    _copy_into(st, new)
    assert ptrs == (st['ref_imgs_info']['imgs'].data_ptr(), st['grasp_info'][0].data_ptr(), st['ref_imgs_info']['bbox3d'].data_ptr())
    assert torch.equal(st['ref_imgs_info']['imgs'], new['ref_imgs_info']['imgs'])
    assert torch.equal(st['grasp_info'][0], new['grasp_info'][0])
    assert np.allclose(st['ref_imgs_info']['bbox3d'].numpy(), new['ref_imgs_info']['bbox3d'])


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(3))

    def forward(self, data):
        return {'y': (self.w * data['x']).sum()}


def test_graph_mode_is_a_no_op_without_cuda():
    """graph=True on a CPU model (or a model without nr_net): the step runs eagerly and gives the eager result."""
    torch.manual_seed(0)
    batch = [{'x': torch.rand(3)} for _ in range(4)]
    res = []
    for graph in (False, True):
        net = _Toy()
        step = TrainStep(net, lr=1e-2, loss_fn=lambda out, data: out['y'], graph=graph)
        res.append(([step(batch) for _ in range(3)], net.w.detach().clone()))
        assert step._g is None
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1])
