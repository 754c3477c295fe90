"""Quick per-kernel CUDA-event timing of the volume path (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graspnerf_b200 import ops
from graspnerf_b200.synth import make_scene
from graspnerf_b200.weights import seed0_weights as golden_weights


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    ops.K2A_IMPL = 'simt' if (len(sys.argv) > 3 and sys.argv[3] == 'simt') else 'tc'
    dev = torch.device('cuda:0')
    sd = golden_weights()
    hw = ops.HeadWeights(sd, 'agg_net.', 'dist_decoder.', dev)
    scs = [make_scene(seed=s) for s in range(B)]
    def stack(k):
        return torch.from_numpy(np.stack([s[k] for s in scs])).to(dev)
    scene = ops.Scene(stack('imgs'), stack('img_feats'), stack('ray_feats'), stack('poses'), stack('Ks'), stack('depth_range'))
    bbox = torch.tensor([s['bbox3d'][0] for s in scs], device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    acc = np.zeros(3)
    for it in range(iters + 3):
        flush.fill_(1.0)
        ev[0].record()
        rec, pt = ops.k1_forward(scene, hw, resolution=40, bbox_min=bbox)
        ev[1].record()
        if ops.K2A_IMPL in ('tc', 'tc3'):
            _, _, _, tok = ops.k2a_forward(rec, pt, hw, scene.depth_range, want_pooled=False, want_tok=True, resolution=40, bbox_min=bbox)
            ev[2].record()
            vol, _ = ops.k2b_forward(None, hw, dn=40, resolution=40, bbox_min=bbox, tok=tok)
        else:
            pooled, _, _ = ops.k2a_forward(rec, pt, hw, scene.depth_range)
            ev[2].record()
            vol, _ = ops.k2b_forward(pooled, hw, dn=40, resolution=40, bbox_min=bbox)
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 3:
            acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    acc /= iters
    print(f'[{ops.K2A_IMPL}] B={B}: K1 {acc[0]*1e3:.1f} us  K2a {acc[1]*1e3:.1f} us  K2b {acc[2]*1e3:.1f} us  total {acc.sum()*1e3:.1f} us '
          f'-> {B/acc.sum()*1e3:.1f} volumes/s; K1 bytes/vol 135876608 -> {135876608*B/acc[0]/1e6:.1f} GB/s')


if __name__ == '__main__':
    main()
