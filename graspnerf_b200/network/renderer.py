"""Drop-in mirror of the reference's model API (src/nr/network/renderer.py:13-335): same class names, ctor cfg handling,
`forward(data)` signature, output-dict keys and `state_dict` keys, so `Trainer` / `GraspNeRFPlanner` can construct
`name2network['grasp_nerf'](cfg)` and load `model_best.pth` unchanged.

What runs where:
  * 2-D encoders, depth-mean head, VGN 3-D conv: PyTorch/cuDNN (out of the CUDA hot path);
  * sample_volume and the RGB head (render): the sm_100a kernels through graspnerf_b200.ops - no torch fallback.
Training: sample_volume has a hand-written first-order backward (gn_k2b_backward -> gn_k2a_backward -> gn_k1_backward,
ops.sample_volume_autograd).  The RGB head (render) trains through ops.ray_features_autograd (K1 / K2a forward kernels,
gn_k2a_backward incl. rgb_fc / gn_k1_backward) plus the small per-ray head and compositing in torch (network/ray_head.py),
which the eikonal loss differentiates twice exactly as the reference does (ibrnet.py:497-504)."""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .encoders import ResUNetLight, VgnConvNet, name2init_net, name2vis_encoder
from .heads import name2agg_net, name2dist_decoder
from . import ray_head
from . import render_ops


class NeuralRayRenderer(nn.Module):
    base_cfg = {                                              # renderer.py:14-47
        'vis_encoder_type': 'default', 'vis_encoder_cfg': {},
        'dist_decoder_type': 'mixture_logistics', 'dist_decoder_cfg': {},
        'agg_net_type': 'default', 'agg_net_cfg': {},
        'use_hierarchical_sampling': False, 'fine_agg_net_cfg': {}, 'fine_dist_decoder_cfg': {},
        'fine_depth_sample_num': 64, 'fine_depth_use_all': False,
        'ray_batch_num': 2048, 'depth_sample_num': 64, 'alpha_value_ground_state': -15,
        'use_dr_prediction': False, 'use_nr_color_for_dr': False, 'use_self_hit_prob': False,
        'use_ray_mask': True, 'ray_mask_view_num': 2, 'ray_mask_point_num': 8,
        'render_depth': False, 'disable_view_dir': False, 'render_rgb': False,
        'init_net_type': 'depth', 'init_net_cfg': {}, 'depth_loss_coords_num': 8192,
    }

    def __init__(self, cfg):
        super().__init__()
        self.cfg = {**self.base_cfg, **cfg}
        if self.cfg['agg_net_type'] != 'neus':
            raise NotImplementedError("only agg_net_type 'neus' (the shipped nrvgn_sdf.yaml) is implemented")
        # configuration values the kernels hard-code: refuse anything else instead of silently computing something different
        for key in ('dist_decoder_cfg', 'fine_dist_decoder_cfg'):
            dcfg = {'bias_val': 0.05, 'use_vis': True, **self.cfg[key]}          # dist_decoder.py:54-58 defaults
            if dcfg['use_vis']:
                raise NotImplementedError(f"{key}.use_vis=True (cdf * vis_decoder, dist_decoder.py:130-131) is not implemented: the shipped "
                                          "nrvgn_sdf.yaml sets use_vis: false for both decoders")
            if float(dcfg['bias_val']) != 0.05:
                raise NotImplementedError(f'{key}.bias_val != 0.05 is not implemented (the K2a kernel adds the 0.05 of dist_decoder.py:56)')
        if self.cfg['disable_view_dir']:
            raise NotImplementedError('disable_view_dir=True (renderer.py:116-117) is not implemented')
        if self.cfg['fine_depth_use_all']:
            raise NotImplementedError('fine_depth_use_all=True (renderer.py:145-146: coarse + fine depths) is not implemented')
        if self.cfg['alpha_value_ground_state'] != -15:
            raise NotImplementedError('alpha_value_ground_state != -15 is not implemented')
        self.vis_encoder = name2vis_encoder[self.cfg['vis_encoder_type']](self.cfg['vis_encoder_cfg'])
        self.dist_decoder = name2dist_decoder[self.cfg['dist_decoder_type']](self.cfg['dist_decoder_cfg'])
        self.image_encoder = ResUNetLight(3, [1, 2, 6, 4], 32, inplanes=16)
        self.init_net = name2init_net[self.cfg['init_net_type']](self.cfg['init_net_cfg'])
        self.agg_net = name2agg_net[self.cfg['agg_net_type']](self.cfg['agg_net_cfg'])
        if self.cfg['use_hierarchical_sampling']:
            self.fine_dist_decoder = name2dist_decoder[self.cfg['dist_decoder_type']](self.cfg['fine_dist_decoder_cfg'])
            self.fine_agg_net = name2agg_net[self.cfg['agg_net_type']](self.cfg['fine_agg_net_cfg'])
        self.agg_net.pair_with(self.dist_decoder, 'agg_net.', 'dist_decoder.')
        if self.cfg['use_hierarchical_sampling']:
            self.fine_agg_net.pair_with(self.fine_dist_decoder, 'fine_agg_net.', 'fine_dist_decoder.')
        self.use_sdf = True
        self._hw = {}
        self._side = None
        self.fused_depth_mean = True                      # False: the torch formulation also in inference (cross-check in the GPU tests)
        self._valid = None
        self.two_stream_encoders = True

    # ------------------------------------------------------------------------------------------------ plumbing
    def _head_weights(self, fine=False):
        key = 'fine' if fine else 'coarse'
        agg, dd = ('fine_agg_net.', 'fine_dist_decoder.') if fine else ('agg_net.', 'dist_decoder.')
        sd = {k: v for k, v in self.named_parameters() if k.startswith(agg) or k.startswith(dd)}
        dev = next(self.parameters()).device
        if key not in self._hw or self._hw[key].device != dev:
            self._hw[key] = ops.HeadWeights(sd, agg, dd, dev)
        else:
            self._hw[key].refresh(sd)
        return self._hw[key]

    @staticmethod
    def _scene(ref_imgs_info):
        # 'imgs_u8' (optional, engine.ForwardEngine / planner): the same images as uint8 RGBA [1,V,H,W,4]; K1 then gathers
        # bytes and divides by 255 itself (bit-identical to gathering the fp32 image, main.py:170)
        imgs = ref_imgs_info.get('imgs_u8', ref_imgs_info['imgs'])
        if imgs.dtype == torch.uint8 and imgs.dim() == 5:
            return ops.Scene(imgs, ref_imgs_info['img_feats'][None], ref_imgs_info['ray_feats'][None], ref_imgs_info['poses'][None],
                             ref_imgs_info['Ks'][None], ref_imgs_info['depth_range'][None])
        return ops.Scene(imgs, ref_imgs_info['img_feats'], ref_imgs_info['ray_feats'],
                         ref_imgs_info['poses'], ref_imgs_info['Ks'], ref_imgs_info['depth_range'])

    # ------------------------------------------------------------------------------------------------ hot path
    def sample_volume(self, ref_imgs_info):
        """renderer.py:164-199 (volume_type ['sdf']): K1 -> K2a -> K2b; returns [1,1,R,R,R]."""
        if any(m != 'sdf' for m in self.cfg['volume_type']):
            raise NotImplementedError("volume_type other than ['sdf'] is not implemented")
        dev = ref_imgs_info['imgs'].device
        # bbox3d is a python list in training (train_dataset.py) and an fp32 tensor [2,3] in inference (main.py:231)
        bbox = ref_imgs_info['bbox3d']
        if torch.is_tensor(bbox) and bbox.device == dev and bbox.dtype == torch.float32:
            bbox_min = bbox.reshape(-1, 3)[0:1]               # already on the device: no pageable H2D copy (which synchronises)
        else:
            bbox_min = torch.as_tensor(bbox[0], dtype=torch.float32).to(dev).reshape(1, 3)
        if torch.is_grad_enabled() and (ref_imgs_info['img_feats'].requires_grad or ref_imgs_info['ray_feats'].requires_grad
                                        or any(p.requires_grad for p in self.agg_net.parameters())):
            named = {k: v for k, v in self.named_parameters() if k.startswith(('agg_net.', 'dist_decoder.'))}
            return ops.sample_volume_autograd(ref_imgs_info['imgs'], ref_imgs_info['img_feats'], ref_imgs_info['ray_feats'],
                                              ref_imgs_info['poses'], ref_imgs_info['Ks'], ref_imgs_info['depth_range'],
                                              bbox_min, named, self.cfg['volume_resolution'])
        scene = self._scene(ref_imgs_info)
        # [1,V,fh,fw,64] channels-last ray|img map: the depth-mean head samples it too (keyed to the tensor it was built from)
        rf = ref_imgs_info['ray_feats']
        ref_imgs_info['_fused_feats'] = (scene.feats, rf.data_ptr(), rf._version)
        # the reference's "!! too low ratio" diagnostic (renderer.py:174-176) without its host synchronisation: K1 counts the
        # valid projections per view into a device word; it is READ when the next call starts (the previous call has long
        # finished by then) or on demand through valid_ratio()
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self._report_valid_ratio()
        if capturing or self._valid is None or self._valid[0].device != dev or self._valid[0].numel() != scene.V:
            # (a captured graph gets a counter of its own: graphs of different engine slots may run concurrently)
            self._valid = [torch.zeros((1, scene.V), dtype=torch.int32, device=dev), 0, False]
        self._valid[0].zero_()
        self._valid[1], self._valid[2] = self.cfg['volume_resolution'] ** 3, not capturing
        return ops.sample_volume(scene, self._head_weights(False), bbox_min, self.cfg['volume_resolution'], valid_count=self._valid[0])

    def valid_ratio(self):
        """Fraction of voxel centres that project into each reference view in the last sample_volume call ([V] tensor on the
        host; synchronises).  renderer.py:174."""
        if self._valid is None:
            return None
        return self._valid[0][0].cpu().float() / float(self._valid[1])

    def _report_valid_ratio(self):
        if self._valid is not None and self._valid[2]:
            self._valid[2] = False
            r = self.valid_ratio()
            if float(r.mean()) < 0.5:
                print("!! too low ratio", r)                     # renderer.py:175-176

    # ------------------------------------------------------------------------------------------------ the reference's stage API
    # renderer.py:62-162 under the reference's names.  render() / sample_volume() use the fused launch sequences in ops
    # directly; these methods expose the same stages one by one (parity tests, diagnostics, external callers).
    def predict_proj_ray_prob(self, prj_dict, ref_imgs_info, que_dists, is_fine):
        """renderer.py:62-78: adds 'alpha', 'vis', 'hit_prob' ([rfn,qn,rn,dn,1]) to prj_dict.  The three decoder MLPs and
        compute_prob run inside K2a; que_dists: [qn,rn,dn] inverse-depth spacings, or an empty tensor (fixed interval)."""
        hw = self._head_weights(is_fine)
        rec, pt, scene = prj_dict['_rec'], prj_dict['_pt'], prj_dict['_scene']
        rfn, qn, rn, dn, _ = prj_dict['mask'].shape
        inv = None if que_dists.numel() == 0 else que_dists.reshape(1, rn * dn).to(rec.device, torch.float32).contiguous()
        _, _, rows = ops.k2a_forward(rec, pt, hw, scene.depth_range, que_dists=inv, dn=dn, debug=True)
        hit = rows[0, :, :, 0].t().reshape(rfn, qn, rn, dn, 1)                               # already x mask (renderer.py:76-77)
        vis = rows[0, :, :, 1].t().reshape(rfn, qn, rn, dn, 1)
        m = prj_dict['mask']
        prj_dict['alpha'] = torch.log(hit / (vis - hit + 1e-5) + 1e-5) * m + (1 - m) * self.cfg['alpha_value_ground_state']
        prj_dict['vis'], prj_dict['hit_prob'], prj_dict['_inv_dists'] = vis, hit, inv
        return prj_dict

    def get_img_feats(self, ref_imgs_info, prj_dict):
        """renderer.py:80-88: the img_feats tap is fused into K1 (one gather serves both feature maps), so project_points_dict
        already produced it; this keeps the reference's call sequence working."""
        if 'img_feats' not in prj_dict:
            raise ValueError('prj_dict must come from graspnerf_b200.network.render_ops.project_points_dict')
        return prj_dict

    def network_rendering(self, prj_dict, que_dir, que_pts, que_depth, is_fine, is_train, is_sdf=False, sdf_only=False):
        """renderer.py:90-108 (is_sdf: the NeuS aggregation net is the only one implemented)."""
        net = self.fine_agg_net if is_fine else self.agg_net
        dists = None
        if que_depth is not None:
            dists = torch.cat([que_depth[..., 1:] - que_depth[..., :-1], torch.full_like(que_depth[..., :1], 1e6)], -1)   # depth2dists
            prj_dict['_que_depth'] = que_depth
        alpha, sdf, colors, grad_err, s = net(prj_dict, que_dir, que_pts, dists, is_train)
        outputs = {'sdf_values': sdf, 'sdf_gradient_error': grad_err, 's': s}
        if sdf_only:
            return outputs
        T = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], -1), -1)[..., :-1]          # render_ops.py:72-80
        outputs.update(alpha_values=alpha, colors_nr=colors, hit_prob_nr=alpha * T)
        outputs['pixel_colors_nr'] = torch.sum(outputs['hit_prob_nr'].unsqueeze(-1) * colors, 2)
        return outputs

    def render_by_depth(self, que_depth, que_imgs_info, ref_imgs_info, is_train, is_fine):
        """renderer.py:110-138, the fused launch sequence (ray set-up -> K1 -> K2a -> K2b -> K3)."""
        que = {k: que_imgs_info[k] for k in ('coords', 'poses', 'Ks', 'depth_range')}
        if torch.is_grad_enabled() and is_train and any(p.requires_grad for p in self.agg_net.parameters()):
            out = ray_head.render_by_depth_autograd(self, ref_imgs_info, que, que_depth, is_fine, is_train)
        else:
            out = ops.render_by_depth(self._scene(ref_imgs_info), self._head_weights(is_fine), que, que_depth,
                                      self.cfg['ray_mask_view_num'], self.cfg['ray_mask_point_num'])
            out['s'] = torch.full((1, 1), self._head_weights(is_fine).variance, device=que_depth.device)
        out.pop('sdf_grad', None)
        if 'imgs' in que_imgs_info:
            out['pixel_colors_gt'] = _bilinear_gt(que_imgs_info['imgs'], que_imgs_info['coords'])
        return out

    def fine_render_impl(self, coarse_render_info, que_imgs_info, ref_imgs_info, is_train):
        """renderer.py:140-150."""
        fd = render_ops.sample_fine_depth(coarse_render_info['depth'], coarse_render_info['hit_prob'], que_imgs_info['depth_range'],
                                          self.cfg['fine_depth_sample_num'], is_train)
        return self.render_by_depth(fd, que_imgs_info, ref_imgs_info, is_train, True)     # the kernel returns them sorted (renderer.py:148)

    def render_impl(self, que_imgs_info, ref_imgs_info, is_train):
        """renderer.py:152-162."""
        que_depth, _ = render_ops.sample_depth(que_imgs_info['depth_range'], que_imgs_info['coords'], self.cfg['depth_sample_num'], False)
        outputs = self.render_by_depth(que_depth, que_imgs_info, ref_imgs_info, is_train, False)
        if self.cfg['use_hierarchical_sampling']:
            fine = self.fine_render_impl({'depth': que_depth, 'hit_prob': outputs['hit_prob_nr']}, que_imgs_info, ref_imgs_info, is_train)
            for k, v in fine.items():
                outputs[k + '_fine'] = v
        return outputs

    def render(self, que_imgs_info, ref_imgs_info, is_train):
        """renderer.py:201-220: chunk the query rays by ray_batch_num, coarse + fine pass per chunk (render_impl 152-162)."""
        coords = que_imgs_info['coords']
        dn, fdn = self.cfg['depth_sample_num'], self.cfg['fine_depth_sample_num']
        autograd = torch.is_grad_enabled() and (ref_imgs_info['img_feats'].requires_grad or ref_imgs_info['ray_feats'].requires_grad
                                                or any(p.requires_grad for p in self.agg_net.parameters()))
        hw_c = hw_f = None
        if not autograd:
            scene = self._scene(ref_imgs_info)
            hw_c = self._head_weights(False)
            hw_f = self._head_weights(True) if self.cfg['use_hierarchical_sampling'] else None
        fine = self.cfg['use_hierarchical_sampling']
        outs = {}
        for s in range(0, coords.shape[1], self.cfg['ray_batch_num']):
            que = {'coords': coords[:, s:s + self.cfg['ray_batch_num']].contiguous(), 'poses': que_imgs_info['poses'],
                   'Ks': que_imgs_info['Ks'], 'depth_range': que_imgs_info['depth_range']}
            rn = que['coords'].shape[1]
            if autograd:      # training: CUDA forward / reverse kernels for the per-(point,view) part, per-ray head in torch
                u = torch.rand(1, rn, fdn, device=coords.device) if (is_train and fine) else None
                res = ray_head.render_rays_autograd(self, ref_imgs_info, que, dn, fdn, u, is_train)
            elif fine:
                # sample_fine_depth uses torch.rand in training (render_ops.py:205), stratified midpoints in eval
                u = torch.rand(1, rn, fdn, device=coords.device) if is_train else None
                res = ops.render_rays(scene, hw_c, hw_f, que, dn, fdn, u)
            else:
                res = ops.render_by_depth(scene, hw_c, que, ops.k3_coarse_depths(que['depth_range'], rn, dn))
            if not autograd:
                res['s'] = torch.full((1, 1), hw_c.variance, device=coords.device)
                if fine:
                    res['s_fine'] = torch.full((1, 1), hw_f.variance, device=coords.device)
            if 'imgs' in que_imgs_info:                     # renderer.py:125-127 (align_corners=True lookup of the GT colours)
                gt = _bilinear_gt(que_imgs_info['imgs'], que['coords'])
                res['pixel_colors_gt'] = gt
                if fine:
                    res['pixel_colors_gt_fine'] = gt
            for k, v in res.items():
                if k in ('depth', 'depth_fine', 'fine_inds', 'sdf_grad', 'sdf_grad_fine'):
                    continue
                outs.setdefault(k, []).append(v)
        return {k: torch.cat(v, 1) for k, v in outs.items()}

    # ------------------------------------------------------------------------------------------------ torch-side heads
    def draw_depth_coords(self, rfn, h, w, device):
        """renderer.py:225-233: depth_loss_coords_num distinct random pixels, the same ones for every reference view ->
        int64 [rfn,num,2].  (Independent of everything else in the forward: engine.ForwardEngine draws them on a side stream.)"""
        idx = torch.randperm(h * w, device=device)[:self.cfg['depth_loss_coords_num']]
        coords = torch.stack([idx // w, idx % w], -1)            # the reference stacks (row, col) of meshgrid(arange(h), arange(w))
        return coords.unsqueeze(0).repeat(rfn, 1, 1)

    def predict_mean_for_depth_loss(self, ref_imgs_info, coords=None):
        """renderer.py:222-266: depth_loss_coords_num random pixels x V -> mean_decoder (+fine).  Inference on the GPU: ONE
        launch (gn_k3_depth_mean: bilinear taps + both decoders); under autograd the torch formulation below.
        coords (optional, [V,num,2] int64): use these pixels instead of drawing new ones (parity tests, engine)."""
        ray_feats, imgs = ref_imgs_info['ray_feats'], ref_imgs_info['imgs']
        rfn, _, h, w = imgs.shape
        if coords is None:
            coords = self.draw_depth_coords(rfn, h, w, imgs.device)
        if ray_feats.is_cuda and not torch.is_grad_enabled() and self.fused_depth_mean:
            fine = self.fine_dist_decoder.mean_decoder if self.cfg['use_hierarchical_sampling'] else None
            fused, ptr, ver = ref_imgs_info.get('_fused_feats', (None, 0, 0))
            if (fused is not None and ptr == ray_feats.data_ptr() and ver == ray_feats._version and fused.shape[0] == 1
                    and tuple(fused.shape[1:4]) == (ray_feats.shape[0],) + tuple(ray_feats.shape[2:])):
                ray_feats = fused[0, ..., :32].permute(0, 3, 1, 2)       # same values, channels-last: one 128-byte run per tap
            m, mf = ops.depth_mean(ray_feats, coords, (h, w), self.dist_decoder.mean_decoder, fine)
            out = {'depth_mean': m[..., 0], 'depth_coords': coords, 'depth_mean_2': m[..., 1]}
            if mf is not None:
                out['depth_mean_fine'], out['depth_mean_fine_2'] = mf[..., 0], mf[..., 1]
            return out
        cf = coords.float()
        grid = torch.stack([cf[..., 0] / (w - 1) * 2 - 1, cf[..., 1] / (h - 1) * 2 - 1], -1).unsqueeze(1)   # ops.py:29-31
        feats = torch.nn.functional.grid_sample(ray_feats, grid, mode='bilinear', padding_mode='border',
                                                align_corners=(ray_feats.shape[-2:] == imgs.shape[-2:])).squeeze(2).permute(0, 2, 1)
        m = self.dist_decoder.predict_mean(feats)
        out = {'depth_mean': m[..., 0], 'depth_coords': coords, 'depth_mean_2': m[..., 1]}
        if self.cfg['use_hierarchical_sampling']:
            mf = self.fine_dist_decoder.predict_mean(feats)
            out['depth_mean_fine'], out['depth_mean_fine_2'] = mf[..., 0], mf[..., 1]
        return out

    def encode(self, ref, src=None, is_train=False):
        """renderer.py:275-279: image_encoder, init_net, vis_encoder -> (img_feats, ray_feats).  In inference on a CUDA device
        the two independent ResUNets (image_encoder and init_net: ~150 small launches each at batch V, latency- not
        throughput-bound) run CONCURRENTLY on two streams (fork / join with events, also under CUDA-graph capture)."""
        imgs = ref['imgs']
        if imgs.is_cuda and not torch.is_grad_enabled() and self.two_stream_encoders:
            cur = torch.cuda.current_stream(imgs.device)
            if self._side is None or self._side.device != imgs.device:
                self._side = torch.cuda.Stream(imgs.device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                ray0 = self.init_net(ref, src, is_train)
            img_feats = self.image_encoder(imgs)
            cur.wait_stream(self._side)
        else:
            img_feats = self.image_encoder(imgs)
            ray0 = self.init_net(ref, src, is_train)
        return img_feats, self.vis_encoder(ray0, img_feats)

    def forward(self, data):
        """renderer.py:268-291."""
        ref = data['ref_imgs_info'].copy()
        que = data['que_imgs_info'].copy()
        is_train = 'eval' not in data
        src = data['src_imgs_info'].copy() if 'src_imgs_info' in data else None
        if 'img_feats' not in ref or 'ray_feats' not in ref:       # (extension) a caller that batches the encoders over several scenes
            ref['img_feats'], ref['ray_feats'] = self.encode(ref, src, is_train)        # passes the scene's slices in (train.TrainStep)
        out = {}
        if self.cfg['render_rgb']:
            out = self.render(que, ref, is_train)
        if self.cfg['sample_volume']:
            out['volume'] = self.sample_volume(ref)
        if (self.cfg['use_depth_loss'] and 'true_depth' in ref) or (not is_train):
            out.update(self.predict_mean_for_depth_loss(ref))
        return out


def _bilinear_gt(imgs, coords):
    """interpolate_feats(imgs, coords, align_corners=True) with the default zero padding (ops.py:14-34, renderer.py:126)."""
    _, _, h, w = imgs.shape
    grid = torch.stack([coords[..., 0] / (w - 1) * 2 - 1, coords[..., 1] / (h - 1) * 2 - 1], -1).unsqueeze(1)
    return torch.nn.functional.grid_sample(imgs, grid, mode='bilinear', padding_mode='zeros', align_corners=True).squeeze(2).permute(0, 2, 1)


class GraspNeRF(nn.Module):
    default_cfg_vgn = {'nr_initial_training_steps': 0, 'freeze_nr_after_init': False}   # renderer.py:294-297

    def __init__(self, cfg):
        super().__init__()
        self.cfg = {**self.default_cfg_vgn, **cfg}
        self.nr_net = NeuralRayRenderer(self.cfg)
        self.vgn_net = VgnConvNet()

    def select(self, out, index):                                                         # renderer.py:305-311
        qual, rot, width = out
        bi = torch.arange(qual.shape[0], device=qual.device)       # (on the device: a host index tensor makes every gather synchronise)
        return (qual[bi, :, index[:, 0], index[:, 1], index[:, 2]].squeeze(), rot[bi, :, index[:, 0], index[:, 1], index[:, 2]],
                width[bi, :, index[:, 0], index[:, 1], index[:, 2]].squeeze())

    def forward(self, data):                                                              # renderer.py:313-331
        out = self.nr_net(data)
        vgn_pred = self.vgn_net(out['volume'])
        out['vgn_pred'] = vgn_pred if 'full_vol' in data else self.select(vgn_pred, data['grasp_info'][0])
        return out


name2network = {'grasp_nerf': GraspNeRF}
