"""Parameter containers of the hot-path heads.  They own the reference's parameters under the reference's key names
(dist_decoder.py:53-98, aggregate_net.py:19-33,87-104, ibrnet.py:373-432, neus.py:6-19) but have NO torch forward for the
heavy part: the math runs in the CUDA kernels (graspnerf_b200.ops).  Initialisation follows the reference (kaiming on
base_fc / vis_fc / vis_fc2 / geometry_fc / rgb_fc / neuray_fc, ibrnet.py:105-109,427-432)."""
import torch
import torch.nn as nn


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
        raise RuntimeError('MixtureLogisticsDistDecoder.forward is fused into the K2a kernel; call NeuralRayRenderer instead')


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

    def forward(self, *a, **k):
        raise RuntimeError('NeusAggregationNet.forward is fused into the K2a/K2b/K3 kernels; call NeuralRayRenderer instead')


name2dist_decoder = {'mixture_logistics': MixtureLogisticsDistDecoder}
name2agg_net = {'neus': NeusAggregationNet}
