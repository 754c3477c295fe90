"""K2a timing aid (development): 16 back-to-back launches over 8 scenes' records between one event pair (see tools/time_k1.py).
GN_LIB_TAG picks a tagged library build (graspnerf_b200/build.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graspnerf_b200 import ops
from graspnerf_b200.synth import make_scene
from graspnerf_b200.weights import seed0_weights


def main():
    dev = torch.device('cuda:0')
    hw = ops.HeadWeights(seed0_weights(), 'agg_net.', 'dist_decoder.', dev)
    items = []
    for s in range(8):
        sc = make_scene(seed=s)
        t = {k: torch.from_numpy(v).to(dev) for k, v in sc.items() if isinstance(v, np.ndarray)}
        scene = ops.Scene(t['imgs'], t['img_feats'], t['ray_feats'], t['poses'], t['Ks'], t['depth_range'])
        bb = torch.tensor([sc['bbox3d'][0]], device=dev)
        rec, pt = ops.k1_forward(scene, hw, resolution=40, bbox_min=bb)
        items.append((scene, bb, rec, pt))
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    for rnd in range(2):
        ts = []
        for rep in range(6):
            for _ in range(3):
                flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for it in range(16):
                scene, bb, rec, pt = items[it % 8]
                ops.k2a_forward(rec, pt, hw, scene.depth_range, want_pooled=False, want_tok=True, resolution=40, bbox_min=bb)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 1:
                ts.append(e0.elapsed_time(e1) * 1e3 / 16)
        print(f'[{os.environ.get("GN_LIB_TAG", "")}] K2a median {np.median(ts):.1f} us  min {np.min(ts):.1f} per 40^3 volume (384 000 rows)')


if __name__ == '__main__':
    main()
