"""The hot-path heads under the reference's class names, parameter names and call signatures (dist_decoder.py:53-153,
aggregate_net.py:19-140, ibrnet.py:373-432, neus.py:6-19).  The modules own the reference's parameters; the heavy math runs
in the CUDA kernels: `NeusAggregationNet.forward(prj_dict, que_dir, que_pts, que_dists, is_train)` launches K1 (dir_diff
with the caller's que_dir) -> K2a -> K2b -> K3 on the record that graspnerf_b200.network.render_ops.project_points_dict
attached to prj_dict.  Initialisation follows the reference (kaiming on base_fc / vis_fc / vis_fc2 / geometry_fc / rgb_fc /
neuray_fc, ibrnet.py:105-109,427-432)."""
import torch
import torch.nn as nn

from .. import ops


def _mlp3(cin, hid, cout, last):
    return nn.Sequential(nn.Linear(cin, hid), nn.ELU(), nn.Linear(hid, hid), nn.ELU(), nn.Linear(hid, cout), last)


class AddBias(nn.Module):
    def __init__(self, val):
        super().__init__()
        self.val = val

    def forward(self, x):
        return x + self.val


class MixtureLogisticsDistDecoder(nn.Module):
    default_cfg = {'feats_dim': 32, 'bias_val': 0.05, 'use_vis': True}

    def __init__(self, cfg):
        super().__init__()
        self.cfg = {**self.default_cfg, **cfg}
        d = self.cfg['feats_dim']
        self.mean_decoder = _mlp3(d, d, 2, nn.Softplus())
        self.var_decoder = nn.Sequential(*_mlp3(d, d, 2, nn.Softplus()), AddBias(self.cfg['bias_val']))
        self.aw_decoder = _mlp3(d, d, 1, nn.Sigmoid())
        if self.cfg['use_vis']:
            self.vis_decoder = _mlp3(d, d, 1, nn.Sigmoid())

    def predict_mean(self, prj_ray_feats):
        """dist_decoder.py:148-150; used by the depth-loss head (renderer.py:244), small and kept in torch."""
        return self.mean_decoder(prj_ray_feats)

    def forward(self, feats):
        """dist_decoder.py:99-107 -> (mean, var, vis=None, aw).  Stand-alone calls evaluate the three 32->32->32->{2,2,1}
        MLPs with the module's own torch layers; the renderer never calls this - predict_proj_ray_prob / render /
        sample_volume run the decoders fused inside K2a (gn_k2a_forward_tc)."""
        return self.mean_decoder(feats), self.var_decoder(feats), None, self.aw_decoder(feats)

    def compute_prob(self, depth, interval, mean, var, vis, aw, is_ref, depth_range):
        """dist_decoder.py:109-142 (use_vis False, is_ref True): mixture-of-logistics visibility / hit probability in
        normalised inverse depth.  Same remark as forward(): the product path has this fused in K2a."""
        near_r, far_r = -1 / depth_range[:, 0], -1 / depth_range[:, 1]
        shape = [-1] + [1] * (depth.dim() - 1)
        d = (-1 / torch.clamp_min(depth, 1e-5) - near_r.reshape(shape)) / (far_r.reshape(shape) - near_r.reshape(shape))   # dist_decoder.py:18-24
        if tuple(interval.shape) != (1, 0):
            half = interval / 2                                                                   # dist_decoder.py:33-38
            first = half[..., 0:1]
            near, far = d - torch.cat([first, half[..., :-1]], -1), d + half
        else:
            near, far = d - 0.005, d + 0.005                                                      # fixed_interval_val 0.01 (121-124, 47-49)
        mix = torch.cat([aw, 1 - aw], -1)
        cdf0 = 0.5 + 0.5 * torch.tanh((near[..., None] - mean) * var)
        cdf1 = 0.5 + 0.5 * torch.tanh((far[..., None] - mean) * var)
        visibility = torch.sum((1 - cdf0) * mix, -1)
        hit_prob = torch.sum((cdf1 - cdf0) * mix, -1)
        return torch.log(hit_prob / (visibility - hit_prob + 1e-5) + 1e-5), visibility, hit_prob


class _Attention(nn.Module):                      # ibrnet.py:52-70 parameters
    def __init__(self):
        super().__init__()
        self.w_qs = nn.Linear(16, 16, bias=False)
        self.w_ks = nn.Linear(16, 16, bias=False)
        self.w_vs = nn.Linear(16, 16, bias=False)
        self.fc = nn.Linear(16, 16, bias=False)
        self.layer_norm = nn.LayerNorm(16, eps=1e-6)


def _kaiming(m):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)


class IBRNetWithNeuRayNeus(nn.Module):            # ibrnet.py:373-432 parameters
    def __init__(self, neuray_in_dim=32, in_feat_ch=32, n_samples=64):
        super().__init__()
        act = nn.ELU(inplace=True)
        self.n_samples = n_samples
        self.ray_dir_fc = nn.Sequential(nn.Linear(4, 16), act, nn.Linear(16, in_feat_ch + 3), act)
        self.base_fc = nn.Sequential(nn.Linear((in_feat_ch + 3) * 5 + neuray_in_dim, 64), act, nn.Linear(64, 32), act)
        self.vis_fc = nn.Sequential(nn.Linear(32, 32), act, nn.Linear(32, 33), act)
        self.vis_fc2 = nn.Sequential(nn.Linear(32, 32), act, nn.Linear(32, 1), nn.Sigmoid())
        self.geometry_fc = nn.Sequential(nn.Linear(32 * 2 + 1 + 21, 64), act, nn.Linear(64, 16), act)
        self.ray_attention = _Attention()
        self.out_geometry_fc = nn.Sequential(nn.Linear(16, 16), nn.Linear(16, 1))
        self.rgb_fc = nn.Sequential(nn.Linear(32 + 1 + 4, 16), act, nn.Linear(16, 8), act, nn.Linear(8, 1))
        self.neuray_fc = nn.Sequential(nn.Linear(neuray_in_dim, 8), act, nn.Linear(8, 1))
        for m in (self.base_fc, self.vis_fc2, self.vis_fc, self.geometry_fc, self.rgb_fc, self.neuray_fc):
            m.apply(_kaiming)


class SingleVarianceNetwork(nn.Module):           # neus.py:6-19
    def __init__(self, init_val, fix_s=-1):
        super().__init__()
        self.register_parameter('variance', nn.Parameter(torch.tensor(init_val)))
        self.variance.requires_grad = False
        self.step, self.fix_s = 0, fix_s

    def set_step(self, step):
        self.step = step


class NeusAggregationNet(nn.Module):
    default_cfg = {'sample_num': 64, 'neuray_dim': 32, 'use_img_feats': False, 'cos_anneal_end_iter': 0, 'init_s': 0.3, 'fix_s': False}

    def __init__(self, cfg):
        super().__init__()
        self.cfg = {**self.default_cfg, **cfg}
        dim = self.cfg['neuray_dim']
        self.prob_embed = nn.Sequential(nn.Linear(2 + 32, dim), nn.ReLU(), nn.Linear(dim, dim))
        self.agg_impl = IBRNetWithNeuRayNeus(dim, n_samples=self.cfg['sample_num'])
        self.deviation_network = SingleVarianceNetwork(self.cfg['init_s'], self.cfg['fix_s'])
        self.step = 0
        self.cos_anneal_ratio = 1.0

    def pair_with(self, dist_decoder, agg_prefix, dd_prefix):
        """The kernels fuse this net with its dist decoder (K2a evaluates both); the renderer tells each agg net which decoder
        it is paired with.  Plain attribute (not a sub-module): the decoder's parameters stay under their own keys."""
        object.__setattr__(self, '_paired', (dist_decoder, agg_prefix, dd_prefix))
        self._hw = None

    def _head_weights(self):
        dd, agg_prefix, dd_prefix = self._paired
        sd = {agg_prefix + k: v for k, v in self.named_parameters()}
        sd.update({dd_prefix + k: v for k, v in dd.named_parameters()})
        dev = next(self.parameters()).device
        if self._hw is None or self._hw.device != dev:
            self._hw = ops.HeadWeights(sd, agg_prefix, dd_prefix, dev)
        else:
            self._hw.refresh(sd)
        return self._hw

    def forward(self, prj_dict, que_dir, que_pts, que_dists, is_train):
        """aggregate_net.py:125-140, evaluation only (training goes through network.ray_head / ops.*_autograd).
        prj_dict: from graspnerf_b200.network.render_ops.project_points_dict (it carries the kernel record);
        que_dir [qn,rn,dn,3]; que_pts [qn,rn,dn,3]; que_dists [qn,rn,dn] metric spacings (depth2dists) or None.
        Returns (alpha, sdf, colors, grad_error, variance) or, with que_dists None, (None, sdf, colors, None, None)."""
        if '_rec' not in prj_dict:
            raise ValueError('prj_dict must come from graspnerf_b200.network.render_ops.project_points_dict (it carries the '
                             'per-(point,view) record the kernels consume)')
        if que_dir is None:
            raise NotImplementedError('que_dir=None (disable_view_dir) is not implemented')
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and is_train:
            raise NotImplementedError('training goes through NeuralRayRenderer.render / sample_volume (autograd nodes), not this call')
        qn, rn, dn, _ = que_pts.shape
        hw = self._head_weights()
        scene, pts = prj_dict['_scene'], prj_dict['_que_pts']
        qd = que_dir[:, :, 0].to(pts.device, torch.float32).contiguous()                   # the same direction for every sample of a ray
        rec, pt = ops.k1_forward(scene, hw, pts=pts, que_dir=qd, dn=dn)                   # dir_diff with the caller's que_dir (aggregate_net.py:11-17)
        inv_dists = prj_dict.get('_inv_dists')
        pooled, colors, _ = ops.k2a_forward(rec, pt, hw, scene.depth_range, que_dists=inv_dists, dn=dn, want_colors=True)
        sdf, grad = ops.k2b_forward(pooled, hw, dn=dn, pts=pts, want_grad=que_dists is not None)
        colors = colors.reshape(qn, rn, dn, 4)[..., :3]
        if que_dists is None:
            return None, sdf.reshape(qn, rn, dn), colors, None, None
        depth = prj_dict.get('_que_depth')
        if depth is None:                                   # differences of the running sum reproduce the spacings
            depth = torch.cumsum(torch.cat([torch.zeros_like(que_dists[..., :1]), que_dists[..., :-1]], -1), -1)
        alpha, _, _, _, eik = ops.k3_composite(sdf.reshape(qn, rn, dn), grad.reshape(qn, rn, dn, 3), torch.cat(
            [colors, torch.zeros_like(colors[..., :1])], -1).contiguous(), qd, depth.contiguous(), ops.inv_s_from(hw.variance), self.cos_anneal_ratio)
        grad_error = (eik.sum(1) / (rn * dn)).reshape(1, 1)
        return alpha, sdf.reshape(qn, rn, dn), colors, grad_error, self.deviation_network.variance.reshape(1, 1)


name2dist_decoder = {'mixture_logistics': MixtureLogisticsDistDecoder}
name2agg_net = {'neus': NeusAggregationNet}
