"""ForwardEngine throughput (development aid): slots x concurrent_slots sweep on the bench's full-forward workload, and a check
that the results do not depend on the mode."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from graspnerf_b200.engine import ForwardEngine, HostScene
from graspnerf_b200.weights import seed0_model
from graspnerf_b200.synth import make_scene


def main():
    dev = torch.device('cuda:0')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = seed0_model().to(dev).eval()
    net.nr_net.cfg['render_rgb'] = False
    pool = []
    for s in range(8):
        pool.append(bench.quantise_images(make_scene(seed=s)))
    hosts = [HostScene(sc['imgs_u8'], None, None, sc['poses'], sc['Ks'], sc['depth_range'], np.asarray(sc['bbox3d'][0], np.float32)) for sc in pool]
    post = dict(tsdf_thres_high=0.0, tsdf_thres_low=-0.85)
    ref = None
    modes = [(3, False, True), (3, True, True), (4, True, True), (6, True, True), (8, True, True), (3, True, False), (6, True, False)]
    for slots, conc, dm in modes:
        eng = ForwardEngine(net, hosts[0], slots=slots, device=dev, post_cfg=post, concurrent_slots=conc, depth_mean=dm)
        res = {}
        for i in range(2 * slots):
            _, fin = eng.submit(hosts[i % 8], tag=i % 8)
            if fin is not None:
                res[fin[0]] = [t.clone() for t in fin[1]]
        for tag, out in eng.drain():
            res[tag] = [t.clone() for t in out]
        if ref is None:
            ref = res
        same = all(all(torch.equal(a, b) for a, b in zip(res[k], ref[k])) for k in res if k in ref)
        n = 240
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            eng.submit(hosts[i % 8], tag=i)
        eng.drain()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f'slots={slots} concurrent={conc} depth_mean={dm} graphed={eng.graphed}: {n / dt:7.1f} volumes/s ({dt / n * 1e3:.3f} ms)  same_as_first={same}', flush=True)
        del eng


if __name__ == '__main__':
    main()
