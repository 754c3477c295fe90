"""Per-kernel CUDA-event timing of the training (forward + backward) volume path on one configs[1] scene."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graspnerf_b200 import ops
from graspnerf_b200.synth import make_scene
from graspnerf_b200.weights import seed0_weights as golden_weights


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device('cuda:0')
    sd = golden_weights()
    hw = ops.HeadWeights(sd, 'agg_net.', 'dist_decoder.', dev)
    scs = [make_scene(seed=s) for s in range(B)]
    st = lambda k: torch.from_numpy(np.stack([s[k] for s in scs])).to(dev)
    scene = ops.Scene(st('imgs'), st('img_feats'), st('ray_feats'), st('poses'), st('Ks'), st('depth_range'))
    bbox = torch.tensor([s['bbox3d'][0] for s in scs], device=dev)
    R = 40
    d_vol = torch.randn(B, 1, R, R, R, device=dev)
    names = ['k1_fwd', 'k2a_fwd(tc)', 'k2b_fwd(full)', 'k2b_bwd', 'k2a_bwd', 'k1_bwd']
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    acc = np.zeros(len(names))
    for it in range(iters + 2):
        torch.cuda.synchronize()
        ev[0].record()
        rec, pt = ops.k1_forward(scene, hw, resolution=R, bbox_min=bbox); ev[1].record()
        pooled, _, _ = ops.k2a_forward(rec, pt, hw, scene.depth_range, impl='tc'); ev[2].record()
        vol, _ = ops.k2b_forward(pooled, hw, dn=R, resolution=R, bbox_min=bbox); ev[3].record()
        d_w = torch.zeros(hw.blob.shape, dtype=torch.float64, device=dev)
        d_pooled = ops.k2b_backward(pooled, hw, d_vol, d_w, dn=R, resolution=R, bbox_min=bbox); ev[4].record()
        d_rec = ops.k2a_backward(rec, pt, hw, scene.depth_range, d_pooled, d_w); ev[5].record()
        d_img, d_ray = ops.k1_backward(scene, hw, d_rec, resolution=R, bbox_min=bbox); ev[6].record()
        torch.cuda.synchronize()
        if it >= 2:
            acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))]
    acc /= iters
    print(f'B={B}: ' + '  '.join(f'{n} {t * 1e3 / B:.0f} us' for n, t in zip(names, acc)) +
          f'  | total {acc.sum() / B:.3f} ms/scene -> {B / acc.sum() * 1e3:.1f} scenes/s (fwd+bwd, volume path)')


if __name__ == '__main__':
    main()
