"""Development aid: lists every host-synchronising torch call of one training step (torch.cuda.set_sync_debug_mode) with the
python frames inside this repo that triggered it."""
import sys, os, traceback, warnings, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from graspnerf_b200.network import name2network, NRVGN_SDF_CFG
from graspnerf_b200.train import TrainStep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
seen = collections.Counter()


def show(message, category, filename, lineno, file=None, line=None):
    if 'synchroniz' not in str(message):
        return
    fr = [f for f in traceback.extract_stack() if f.filename.startswith(ROOT) and 'find_syncs' not in f.filename]
    key = ' <- '.join(f'{os.path.relpath(f.filename, ROOT)}:{f.lineno}' for f in reversed(fr[-4:]))
    if len(fr) <= 2:
        allf = [f for f in traceback.extract_stack() if 'find_syncs' not in f.filename and 'warnings' not in f.filename]
        key += ' || ' + ' <- '.join(f'{os.path.basename(f.filename)}:{f.lineno}:{f.name}' for f in reversed(allf[-7:]))
    seen[key] += 1


def main():
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    net = name2network['grasp_nerf'](dict(NRVGN_SDF_CFG)).to(dev).train()
    step = TrainStep(net, lr=1e-4)
    batch = [bench.make_train_data(i, dev) for i in range(2)]
    step(batch)
    step(batch)
    torch.cuda.synchronize()
    warnings.simplefilter('always')
    warnings.showwarning = show
    torch.cuda.set_sync_debug_mode('warn')
    step(batch)
    torch.cuda.set_sync_debug_mode('default')
    for k, n in seen.most_common():
        print(n, k)


if __name__ == '__main__':
    main()
