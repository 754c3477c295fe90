"""K1 timing aid (development): variants selected by environment (GN_K1_BULK=0/1, GN_K1_TEXPF=0/1; GN_LIB_TAG picks a
tagged library build).  Each sample = 16 back-to-back launches cycling through 8 scenes behind an L2-flushing fill (event
timestamps on this GPU tick at 4.096 us, so single launches cannot be timed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graspnerf_b200 import ops
from graspnerf_b200.synth import make_scene
from graspnerf_b200.weights import seed0_weights


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    cfgs = [dict(kv.split('=') for kv in a.split(',')) for a in sys.argv[2:]] or [dict(GN_K1_BULK='0'), dict(GN_K1_BULK='1'), dict(GN_K1_BULK='0'), dict(GN_K1_BULK='1')]
    dev = torch.device('cuda:0')
    hw = ops.HeadWeights(seed0_weights(), 'agg_net.', 'dist_decoder.', dev)
    scenes = []
    for s in range(8):
        sc = make_scene(seed=s)
        t = {k: torch.from_numpy(v).to(dev) for k, v in sc.items() if isinstance(v, np.ndarray)}
        scenes.append((ops.Scene(t['imgs'], t['img_feats'], t['ray_feats'], t['poses'], t['Ks'], t['depth_range']),
                       torch.tensor([sc['bbox3d'][0]], device=dev)))
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    for cfg in cfgs:
        os.environ.update(cfg)
        impl = ' '.join(f'{k}={v}' for k, v in cfg.items())
        ts = []
        for rep in range(iters // 8 + 2):
            for _ in range(3):
                flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for it in range(16):          # 16 launches over 8 different scenes: inputs (8 x 28 MB) and records (8 x 110 MB) cycle past L2
                sc, bb = scenes[it % 8]
                rec, pt = ops.k1_forward(sc, hw, resolution=40, bbox_min=bb)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3 / 16)
        ts = np.array(ts)
        print(f'[{os.environ.get("GN_LIB_TAG", "")}] {impl}: K1 median {np.median(ts):.1f} us  min {ts.min():.1f}  mean {ts.mean():.1f} '
              f'-> {135876608 / np.median(ts) / 1e3:.0f} GB/s (own bytes 135.9 MB)')


if __name__ == '__main__':
    main()
