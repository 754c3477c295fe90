"""Development aid: encoders of one scene in a CUDA graph under different K7 split-K policies (ops.K7_SPLIT_POLICY)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graspnerf_b200 import ops
from graspnerf_b200.synth import make_scene
from graspnerf_b200.weights import seed0_model
from profile_forward import timeit

dev = torch.device('cuda:0')
torch.backends.cudnn.allow_tf32 = False
net = seed0_model().to(dev).eval()
nr = net.nr_net
imgs = torch.from_numpy(make_scene(seed=0)['imgs']).to(dev)
policies = {'current': ((64, 16, 4), (160, 16, 2)), 'none': (), 's8/4': ((64, 16, 8), (160, 16, 4)), 's8/2': ((64, 16, 8), (160, 16, 2)),
            's4/4': ((64, 16, 4), (160, 16, 4)), 's4/2/2': ((64, 16, 4), (160, 16, 2), (600, 8, 2)), 's2/2': ((64, 16, 2), (160, 16, 2)), 's4/1': ((64, 16, 4),)}
ref = None
with torch.no_grad():
    for two in (False, True):
        nr.two_stream_encoders = two
        for name, pol in policies.items():
            ops.K7_SPLIT_POLICY = pol
            g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                nr.encode({'imgs': imgs}, None, False); torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    out = nr.encode({'imgs': imgs}, None, False)
            t = timeit(lambda: g.replay())
            if ref is None:
                ref = [o.clone() for o in out]
            err = max(float((a - b).abs().max()) for a, b in zip(out, ref))
            print(f'two_stream={two} policy {name:8s}: {t:.3f} ms   max|diff| vs first {err:.2e}', flush=True)
            del g
