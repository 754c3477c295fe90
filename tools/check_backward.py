"""Development aid: per-stage and end-to-end comparison of the CUDA backward kernels with torch autograd through the
oracle (CPU) and with the reference's own gradients (tests/golden/volume_grad_small_v4.npz).  Prints one line per tensor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from graspnerf_b200 import ops
from graspnerf_b200.synth import make_scene
from graspnerf_b200.weights import unpack_blob_grad
from oracle import nr_oracle as O
from tests.helpers import golden_weights, load_golden, oracle_volume_grads
from tests.golden.cases import VOLUME_CASES


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)), float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
    dev = torch.device('cuda:0')
    g = load_golden('volume_grad_small_v4.npz')
    sd = {k: v for k, v in golden_weights().items() if k.startswith(('agg_net.', 'dist_decoder.'))}
    sc = make_scene(**VOLUME_CASES['small_v4'])
    sct = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in sc.items()}
    R = 40
    # ---------------- end to end
    params = {k: torch.nn.Parameter(v.clone().to(dev)) for k, v in sd.items()}
    imgf = sct['img_feats'].to(dev).requires_grad_(True)
    rayf = sct['ray_feats'].to(dev).requires_grad_(True)
    bbox = torch.tensor([sc['bbox3d'][0]], device=dev)
    vol = ops.sample_volume_autograd(sct['imgs'].to(dev), imgf, rayf, sct['poses'].to(dev), sct['Ks'].to(dev),
                                     sct['depth_range'].to(dev), bbox, params, R)
    G = torch.from_numpy(g['G']).to(dev)
    loss = (vol * G).sum()
    loss.backward()
    torch.cuda.synchronize()
    print('loss ours %.6f  reference %.6f' % (loss.item(), float(g['loss'])))
    g64 = load_golden('volume_grad64_small_v4.npz')
    print('vs fp64 oracle:  d_img ours %.2e ref32 %.2e | d_ray ours %.2e ref32 %.2e' % (
        rel(imgf.grad.cpu(), g64['d_img_feats'])[0], rel(g['d_img_feats'], g64['d_img_feats'])[0],
        rel(rayf.grad.cpu(), g64['d_ray_feats'])[0], rel(g['d_ray_feats'], g64['d_ray_feats'])[0]))
    print('d_img_feats  max/l2 rel vs reference', rel(imgf.grad.cpu(), g['d_img_feats']))
    print('d_ray_feats  max/l2 rel vs reference', rel(rayf.grad.cpu(), g['d_ray_feats']))
    worst = 0.0
    for k in sorted(params):
        gk = 'dw/' + k
        if gk not in g or 'rgb_fc' in k:
            continue
        if params[k].grad is None:
            print('%-60s MISSING' % k); continue
        r = rel(params[k].grad.cpu(), g[gk])
        worst = max(worst, r[0])
        print('%-60s max %.2e l2 %.2e | vs fp64: ours %.2e ref32 %.2e' % (k, r[0], r[1], rel(params[k].grad.cpu(), g64[gk])[0], rel(g[gk], g64[gk])[0]))
    print('worst parameter-gradient max-rel error: %.2e' % worst)
    os.makedirs('gpurun_out', exist_ok=True)
    np.savez_compressed('gpurun_out/ours_grads.npz', d_img=imgf.grad.cpu().numpy(), d_ray=rayf.grad.cpu().numpy(), **{k: v.grad.cpu().numpy() for k, v in params.items() if v.grad is not None})

    # ---------------- stage: K2b backward alone (torch head from pooled on the CPU)
    with torch.no_grad():
        hw = ops.HeadWeights(sd, 'agg_net.', 'dist_decoder.', dev)
        scene = ops.Scene(sct['imgs'].to(dev), sct['img_feats'].to(dev), sct['ray_feats'].to(dev), sct['poses'].to(dev),
                          sct['Ks'].to(dev), sct['depth_range'].to(dev))
        rec, pt = ops.k1_forward(scene, hw, resolution=R, bbox_min=bbox)
        pooled, _, _ = ops.k2a_forward(rec, pt, hw, scene.depth_range, impl='simt')
        d_w = torch.zeros(hw.blob.shape, dtype=torch.float64, device=dev)
        d_pooled = ops.k2b_backward(pooled, hw, G, d_w, dn=R, resolution=R, bbox_min=bbox)
        torch.cuda.synchronize()
    A = 'agg_net.agg_impl.'
    sdc = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(A + 'geometry_fc') or k.startswith(A + 'ray_attention') or k.startswith(A + 'out_geometry')}
    pc = pooled[0].cpu()
    pin = pc[:, :65].clone().requires_grad_(True)
    nvalid = pc[:, 65]
    pts = O.volume_query_points(sc['bbox3d'][0], R, 0.3, torch.float32).reshape(R * R, R, 3)
    gfeat = torch.cat([pin.reshape(R * R, R, 65), O.embed_points(pts)], -1)
    gg = F.elu(O._lin(sdc, A + 'geometry_fc.2', F.elu(O._lin(sdc, A + 'geometry_fc.0', gfeat))))
    gg = gg + O.positional_table(R)[None]
    gg = O.ray_attention(sdc, A, gg, (nvalid.reshape(R * R, R) > 1).float())
    sdf = O._lin(sdc, A + 'out_geometry_fc.1', O._lin(sdc, A + 'out_geometry_fc.0', gg)).clip(-1, 1)[..., 0]
    sdf = sdf.masked_fill(nvalid.reshape(R * R, R) < 1, 1.0)
    volc = sdf.reshape(1, 1, R, R, R).flip(-1)
    (volc * torch.from_numpy(g['G'])).sum().backward()
    print('[K2b] d_pooled', rel(d_pooled[0, :, :65].cpu(), pin.grad))
    gk = unpack_blob_grad(d_w.float().cpu())
    for k in sorted(sdc):
        print('[K2b] %-55s' % k, rel(gk[k], sdc[k].grad))

    # ---------------- stage: K1 backward alone
    rng = np.random.default_rng(5)
    d_rec = torch.from_numpy(rng.standard_normal((1, R ** 3, scene.V, 64)).astype(np.float32)).to(dev)
    d_img, d_ray = ops.k1_backward(scene, hw, d_rec, resolution=R, bbox_min=bbox)
    torch.cuda.synchronize()
    imc = sct['img_feats'].clone().requires_grad_(True)
    rac = sct['ray_feats'].clone().requires_grad_(True)
    sc2 = dict(sct); sc2['img_feats'] = imc; sc2['ray_feats'] = rac
    pts_all = O.volume_query_points(sc['bbox3d'][0], R, 0.3, torch.float32).reshape(-1, 3)
    r = O.project_and_sample(sc2, pts_all)
    dr = d_rec[0].cpu()
    ((r['ray_feats'] * dr[..., :32]).sum() + (r['img_feats'] * dr[..., 32:]).sum()).backward()
    print('[K1] d_ray_feats', rel(d_ray[0].permute(0, 3, 1, 2).cpu(), rac.grad))
    print('[K1] d_img_feats', rel(d_img[0].permute(0, 3, 1, 2).cpu(), imc.grad))


if __name__ == '__main__':
    main()
