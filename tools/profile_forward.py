"""Development aid: where does the time of the planner's whole network call go?  Prints the top CUDA kernels of one eager
GraspNeRF.forward (eval, render_rgb off) of the mirror and times the 2-D encoders under a few settings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graspnerf_b200.synth import make_scene
from graspnerf_b200.weights import seed0_model


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device('cuda:0')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = seed0_model().to(dev).eval()
    net.nr_net.cfg['render_rgb'] = False
    nr = net.nr_net
    sc = make_scene(seed=0)
    imgs = torch.from_numpy(sc['imgs']).to(dev)
    ref = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in sc.items() if k not in ('img_feats', 'ray_feats')}

    def encoders(x):
        return nr.encode({'imgs': x}, None, False)
    with torch.no_grad():
        from graspnerf_b200.network import encoders as E
        for fused, tc in ((False, False), (True, False), (True, True)):
            for two in (False, True):
                E.FUSED, E.TC_CONV, nr.two_stream_encoders = fused, tc, two
                print(f'encoders eager fp32 fused={fused} tc_conv={tc} two_stream={two}: {timeit(lambda: encoders(imgs)):.3f} ms')
                g0 = torch.cuda.CUDAGraph(); s0 = torch.cuda.Stream()
                with torch.cuda.stream(s0):
                    encoders(imgs); torch.cuda.synchronize()
                    with torch.cuda.graph(g0, stream=s0):
                        o0 = encoders(imgs)
                print(f'encoders GRAPH fp32 fused={fused} tc_conv={tc} two_stream={two}: {timeit(lambda: g0.replay()):.3f} ms')
                del g0
        print(f'  image_encoder only              : {timeit(lambda: nr.image_encoder(imgs)):.3f} ms')
        print(f'  init_net only                   : {timeit(lambda: nr.init_net({"imgs": imgs}, None, False)):.3f} ms')
        f = nr.image_encoder(imgs); r = nr.init_net({'imgs': imgs}, None, False)
        print(f'  vis_encoder only                : {timeit(lambda: nr.vis_encoder(r, f)):.3f} ms')
        vol = torch.rand(1, 1, 40, 40, 40, device=dev) * 2 - 1
        print(f'vgn_net                           : {timeit(lambda: net.vgn_net(vol)):.3f} ms')
        ref2 = dict(ref); ref2['img_feats'], ref2['ray_feats'] = encoders(imgs)
        print(f'sample_volume (eager, 3 launches) : {timeit(lambda: nr.sample_volume(ref2)):.3f} ms')
        print(f'depth-mean head                   : {timeit(lambda: nr.predict_mean_for_depth_loss(ref2)):.3f} ms')
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            encoders(imgs)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                out = encoders(imgs)
        print(f'encoders CUDA graph fp32 NCHW     : {timeit(lambda: g.replay()):.3f} ms')
        torch.backends.cudnn.allow_tf32 = True
        print(f'encoders eager TF32 NCHW          : {timeit(lambda: encoders(imgs)):.3f} ms')
        torch.backends.cudnn.allow_tf32 = False
        from torch.profiler import profile, ProfilerActivity
        for m in (nr.image_encoder, nr.init_net, nr.vis_encoder):
            m.to(memory_format=torch.contiguous_format)
        torch.backends.cudnn.benchmark = True
        print(f'encoders eager fp32 NCHW cudnn.benchmark : {timeit(lambda: encoders(imgs)):.3f} ms')
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(3):
                encoders(imgs)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=70))


if __name__ == '__main__':
    main()
