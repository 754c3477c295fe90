"""Development aid for ncu launch lists: a few eager (un-graphed) full forwards of the mirror (GraspNeRF.forward, eval, render_rgb off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graspnerf_b200.synth import make_scene, make_query
from graspnerf_b200.weights import seed0_model

dev = torch.device('cuda:0')
torch.backends.cudnn.allow_tf32 = False
net = seed0_model().to(dev).eval()
net.nr_net.cfg['render_rgb'] = False
net.nr_net.two_stream_encoders = False          # serialised: per-kernel times without overlap
sc = make_scene(seed=0)
ref = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in sc.items() if k not in ('img_feats', 'ray_feats')}
q = {k: torch.from_numpy(v).to(dev) for k, v in make_query(sc, 16, 7).items() if isinstance(v, np.ndarray)}
data = {'step': 0, 'eval': True, 'full_vol': True, 'ref_imgs_info': ref, 'que_imgs_info': q, 'src_imgs_info': ref}
with torch.no_grad():
    for _ in range(8):
        out = net(dict(data))
torch.cuda.synchronize()
print(float(out['volume'].mean()))
