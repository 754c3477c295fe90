"""Development aid: which backward node synchronises with the host?  (sync debug mode 'error' + anomaly mode: the error names
the node and prints the forward call that created it)"""
import sys, os, traceback, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from graspnerf_b200.network import name2network, NRVGN_SDF_CFG
from graspnerf_b200.train import TrainStep, _static_like

dev = torch.device('cuda:0')
torch.manual_seed(0)
net = name2network['grasp_nerf'](dict(NRVGN_SDF_CFG)).to(dev).train()
step = TrainStep(net, lr=1e-4)
batch = [_static_like(bench.make_train_data(i, dev), dev) for i in range(2)]
step(batch); step(batch)
torch.cuda.synchronize()
with torch.autograd.set_detect_anomaly(True, check_nan=False):
    loss = step._group_loss(batch)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode('error')
    try:
        loss.backward()
    except Exception as e:
        print('ERROR:', str(e)[:600])
    torch.cuda.set_sync_debug_mode('default')
