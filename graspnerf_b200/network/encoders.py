"""2-D feature encoders and the VGN 3-D head (the steps right before / after the hot path).  The modules
own every parameter of the reference checkpoint under the reference's key names; architectures follow
src/nr/network/ops.py:78-230, init_net.py:8-35, vis_encoder.py:6-21 and src/gd/networks.py:39-97 (key/shape table pinned by
tests/golden/state_dict_keys.json).

Two execution paths per module, same parameters:
  * `forward`       plain PyTorch/cuDNN, differentiable: training;
  * `forward_fused` inference on CUDA tensors: convolutions in cuDNN fp32, everything between two convolutions (reflection
                    pad, InstanceNorm, activation, residual add, bilinear x2 upsampling) in ONE launch of csrc/k6_encoder_fused.cu
                    per layer; the VGN head entirely in csrc/k5_vgn_conv.cu.
`use_fused(x)` picks the path."""
import torch
import torch.nn as nn
import torch.nn.functional as F

FUSED = True          # module-level switch for the fused inference path (tests compare both paths)
TC_CONV = True        # fused path: convolutions on tcgen05 (csrc/k7_conv_tc.cu) instead of cuDNN fp32


def use_fused(x, module):
    return FUSED and x.is_cuda and x.dtype == torch.float32 and not (
        torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters())))


def _conv0(xp, conv, allow_split=False):
    """A convolution whose (reflection) padding has already been applied by the producer of `xp`: implicit GEMM on tcgen05
    (K7) when the output-channel count fits its tiles, cuDNN otherwise."""
    if TC_CONV and xp.is_cuda and conv.weight.shape[0] <= 128 and (conv.weight.shape[0] + 15) // 16 * 16 in (16, 32, 48, 64, 128):
        from .. import ops
        return ops.conv2d_tc(xp.contiguous(), conv, allow_split=allow_split)
    return F.conv2d(xp, conv.weight, conv.bias, conv.stride, 0)


def _inorm(c):
    return nn.InstanceNorm2d(c, track_running_stats=False, affine=True)


def _c3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 3, stride, 1, bias=False, padding_mode='reflect')


def _c1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 1, stride, bias=False, padding_mode='reflect')


class BasicBlock(nn.Module):                      # ops.py:86-125
    def __init__(self, cin, cout, stride=1, downsample=None):
        super().__init__()
        self.conv1, self.bn1 = _c3(cin, cout, stride), _inorm(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2, self.bn2 = _c3(cout, cout), _inorm(cout)
        self.downsample = downsample

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu(y + (x if self.downsample is None else self.downsample(x)))

    def forward_fused(self, xp, xu, want_padded=True, want_unpadded=False):
        """xp: block input reflection-padded by 1; xu: the same tensor un-padded (only read by a 1x1 downsample branch).
        -> (output padded by 1 or None, output un-padded or None)."""
        from .. import ops
        a_p, _ = ops.norm_act_pad(_conv0(xp, self.conv1, True), self.bn1, 'relu', pad=1)      # split-K partials are summed by the norm stage
        raw2 = _conv0(a_p, self.conv2, True)
        if self.downsample is None:          # identity residual = interior of the padded input
            return ops.norm_act_pad(raw2, self.bn2, 'relu', pad=1, res=xp, res_pad=1, want_padded=want_padded, want_unpadded=want_unpadded)
        rawd = _conv0(xu, self.downsample[0])
        return ops.norm_act_pad(raw2, self.bn2, 'relu', pad=1, res=rawd, res_norm=self.downsample[1],
                                want_padded=want_padded, want_unpadded=want_unpadded)


class ConvNormELU(nn.Module):                     # ops.py `conv` 127-140
    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, padding_mode='reflect')
        self.bn = _inorm(cout)

    def forward(self, x):
        return F.elu(self.bn(self.conv(x)))

    def forward_fused(self, xp, pad=0, want_padded=False, want_unpadded=True):
        """xp: input already reflection-padded for self.conv."""
        from .. import ops
        return ops.norm_act_pad(_conv0(xp, self.conv), self.bn, 'elu', pad=pad, want_padded=want_padded, want_unpadded=want_unpadded)


class UpConv(nn.Module):                          # ops.py `upconv` 142-150
    def __init__(self, cin, cout, k, scale):
        super().__init__()
        self.scale = scale
        self.conv = ConvNormELU(cin, cout, k, 1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=self.scale, mode='bilinear', align_corners=True))

    def forward_fused(self, xu):
        from .. import ops
        assert self.scale == 2
        return self.conv.forward_fused(ops.upsample2x_pad(xu.contiguous(), pad=(self.conv.conv.kernel_size[0] - 1) // 2))[1]


class ResUNetLight(nn.Module):                    # ops.py:150-230
    def __init__(self, in_dim=3, layers=(2, 3, 6, 3), out_dim=32, inplanes=32):
        super().__init__()
        self.inplanes = inplanes
        self.conv1 = nn.Conv2d(in_dim, inplanes, 7, 2, 3, bias=False, padding_mode='reflect')
        self.bn1 = _inorm(inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._stage(32, layers[0])
        self.layer2 = self._stage(64, layers[1])
        self.layer3 = self._stage(128, layers[2])
        self.upconv3 = UpConv(128, 64, 3, 2)
        self.iconv3 = ConvNormELU(128, 64, 3, 1)
        self.upconv2 = UpConv(64, 32, 3, 2)
        self.iconv2 = ConvNormELU(64, 32, 3, 1)
        self.out_conv = nn.Conv2d(32, out_dim, 1, 1)

    def _stage(self, planes, blocks):
        down = nn.Sequential(_c1(self.inplanes, planes, 2), _inorm(planes))     # stride 2 on every stage
        mods = [BasicBlock(self.inplanes, planes, 2, down)]
        self.inplanes = planes
        mods += [BasicBlock(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*mods)

    @staticmethod
    def _skip(skip, x):
        dy, dx = x.shape[2] - skip.shape[2], x.shape[3] - skip.shape[3]
        skip = F.pad(skip, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
        return torch.cat([x, skip], 1)

    def forward(self, x):
        if use_fused(x, self):
            return self.forward_fused(x)
        x = self.relu(self.bn1(self.conv1(x)))
        x1 = self.layer1(x)
        x2 = self.layer2(x1)
        x3 = self.layer3(x2)
        x = self.iconv3(self._skip(x2, self.upconv3(x3)))
        x = self.iconv2(self._skip(x1, self.upconv2(x)))
        return self.out_conv(x)

    @staticmethod
    def _stage_fused(stage, xp, xu, last_padded, last_unpadded):
        n = len(stage)
        for i, blk in enumerate(stage):
            last = i == n - 1
            xp, xu = blk.forward_fused(xp, xu, want_padded=(not last) or last_padded, want_unpadded=last and last_unpadded)
        return xp, xu

    def forward_fused(self, x):
        """Same math as forward(), inference only: every tensor between two convolutions is produced by one K6 launch,
        already reflection-padded for its consumer."""
        from .. import ops
        xin, _ = ops.norm_act_pad(x.contiguous(), None, None, pad=3)                                  # reflection pad for the 7x7 stem
        x0p, x0u = ops.norm_act_pad(_conv0(xin, self.conv1), self.bn1, 'relu', pad=1, want_unpadded=True)
        x1p, x1u = self._stage_fused(self.layer1, x0p, x0u, True, True)
        x2p, x2u = self._stage_fused(self.layer2, x1p, x1u, True, True)
        _, x3u = self._stage_fused(self.layer3, x2p, x2u, False, True)
        y = self._skip(x2u, self.upconv3.forward_fused(x3u))
        yp, _ = ops.norm_act_pad(y.contiguous(), None, None, pad=1)
        y = self.iconv3.forward_fused(yp)[1]
        y = self._skip(x1u, self.upconv2.forward_fused(y))
        yp, _ = ops.norm_act_pad(y.contiguous(), None, None, pad=1)
        y = self.iconv2.forward_fused(yp)[1]
        return _conv0(y, self.out_conv)                    # 1x1 + bias (as a module call: cuDNN + its index pre-computation kernel)


class ResidualBlock(nn.Module):                   # ops.py:43-76 (use_norm branch)
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Sequential(_inorm(cin), nn.ReLU(True), nn.Conv2d(cin, cout, 3, 1, 1, bias=False, padding_mode='reflect'),
                                  _inorm(cout), nn.ReLU(True), nn.Conv2d(cout, cout, 3, 1, 1, bias=False, padding_mode='reflect'))
        self.short_cut = None if cin == cout else nn.Conv2d(cin, cout, 1, 1)

    def forward(self, x):
        y = self.conv(x)
        return y + (x if self.short_cut is None else self.short_cut(x))

    def forward_fused(self, x):
        """pre-activation residual block: IN -> ReLU -> conv3x3 -> IN -> ReLU -> conv3x3, + x."""
        from .. import ops
        t, _ = ops.norm_act_pad(x.contiguous(), self.conv[0], 'relu', pad=1)
        t, _ = ops.norm_act_pad(_conv0(t, self.conv[2]), self.conv[3], 'relu', pad=1)
        return _conv0(t, self.conv[5]) + (x if self.short_cut is None else self.short_cut(x))


class CostVolumeInitNet(nn.Module):               # init_net.py:8-35 (no cost volume despite the name)
    default_cfg = {'cost_volume_sn': 64}

    def __init__(self, cfg):
        super().__init__()
        self.cfg = {**self.default_cfg, **cfg}
        self.register_buffer('imagenet_mean', torch.tensor([0.485, 0.456, 0.406])[None, :, None, None])
        self.register_buffer('imagenet_std', torch.tensor([0.229, 0.224, 0.225])[None, :, None, None])
        self.res_net = ResUNetLight(out_dim=32)
        self.out_conv = nn.Sequential(_c3(32, 32), ResidualBlock(32, 32), _c1(32, 32))

    def forward(self, ref_imgs_info, src_imgs_info, is_train):
        x = self.res_net(ref_imgs_info['imgs'])
        if use_fused(x, self):
            from .. import ops
            xp, _ = ops.norm_act_pad(x.contiguous(), None, None, pad=1)
            # (the 1x1 module has padding_mode='reflect' with padding 0: nn.Conv2d would still launch a pad kernel that copies)
            return _conv0(self.out_conv[1].forward_fused(_conv0(xp, self.out_conv[0])), self.out_conv[2])
        return self.out_conv(x)


class DefaultVisEncoder(nn.Module):               # vis_encoder.py:6-21
    def __init__(self, cfg):
        super().__init__()
        self.cfg = dict(cfg)
        self.out_conv = nn.Sequential(_c3(64, 32), ResidualBlock(32, 32), ResidualBlock(32, 32), _c1(32, 32))

    def forward(self, ray_feats, imgs_feats):
        x = torch.cat([imgs_feats, ray_feats], 1)
        if use_fused(x, self):
            from .. import ops
            xp, _ = ops.norm_act_pad(x, None, None, pad=1)
            y = self.out_conv[1].forward_fused(_conv0(xp, self.out_conv[0]))
            return _conv0(self.out_conv[2].forward_fused(y), self.out_conv[3])
        return self.out_conv(x)


name2init_net = {'cost_volume': CostVolumeInitNet}
name2vis_encoder = {'default': DefaultVisEncoder}


class _VgnEncoder(nn.Module):                     # gd/networks.py:57-74
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv3d(1, 16, 5, stride=2, padding=2)
        self.conv2 = nn.Conv3d(16, 32, 3, stride=2, padding=1)
        self.conv3 = nn.Conv3d(32, 64, 3, stride=2, padding=1)

    def forward(self, x):
        return F.relu(self.conv3(F.relu(self.conv2(F.relu(self.conv1(x))))))


class _VgnDecoder(nn.Module):                     # gd/networks.py:77-97
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv3d(64, 64, 3, padding=1)
        self.conv2 = nn.Conv3d(64, 32, 3, padding=1)
        self.conv3 = nn.Conv3d(32, 16, 5, padding=2)

    def forward(self, x):
        # the reference hard-codes the grids 10 / 20 / 40 (networks.py:88-96) = x2 per stage for its 40^3 volume; written as
        # x2 here so that the 80^3 volumes of BASELINE configs[4] work too
        n = x.shape[-1]
        x = F.interpolate(F.relu(self.conv1(x)), 2 * n)
        x = F.interpolate(F.relu(self.conv2(x)), 4 * n)
        return F.interpolate(F.relu(self.conv3(x)), 8 * n)


class VgnConvNet(nn.Module):                      # gd/networks.py:39-54
    def __init__(self):
        super().__init__()
        self.encoder, self.decoder = _VgnEncoder(), _VgnDecoder()
        self.conv_qual = nn.Conv3d(16, 1, 5, padding=2)
        self.conv_rot = nn.Conv3d(16, 4, 5, padding=2)
        self.conv_width = nn.Conv3d(16, 1, 5, padding=2)

    def forward(self, x, out=None):
        """networks.py:47-54.  Inference on a CUDA volume: seven direct-convolution launches (csrc/k5_vgn_conv.cu, upsampling
        folded into the weights).  With autograd recording (training) the same layers run through torch / cuDNN."""
        if x.is_cuda and not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))):
            from .. import ops
            if getattr(self, '_vw', None) is None:
                object.__setattr__(self, '_vw', ops.VgnWeights(self))
            return ops.vgn_forward(x, self._vw, out=out)
        return self.forward_torch(x)

    def forward_torch(self, x):
        x = self.decoder(self.encoder(x))
        return torch.sigmoid(self.conv_qual(x)), F.normalize(self.conv_rot(x), dim=1), self.conv_width(x)
