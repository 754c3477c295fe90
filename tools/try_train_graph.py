"""Development aid: TrainStep(graph=True) on the bench's training scenes - does the capture go through, does a replay on NEW
inputs reproduce the eager loss, and what does a step cost either way."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import copy
import torch
import bench
from graspnerf_b200.network import name2network, NRVGN_SDF_CFG
from graspnerf_b200.train import TrainStep


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    net = name2network['grasp_nerf'](dict(NRVGN_SDF_CFG)).to(dev).train()
    net2 = copy.deepcopy(net)
    A = [bench.make_train_data(i, dev) for i in range(nb)]
    Bt = [bench.make_train_data(100 + i, dev) for i in range(nb)]
    eager, graphed = TrainStep(net2, lr=1e-4), TrainStep(net, lr=1e-4, graph=True)
    for name, step in (('eager', eager), ('graph', graphed)):
        torch.manual_seed(1)
        ls = [step(A), step(Bt), step(A), step(Bt)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            ls.append(step(A)); ls.append(step(Bt))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        print(f'{name}: {nb / dt:6.1f} scenes/s ({dt * 1e3:.1f} ms/step of {nb} scenes)  losses ' + ' '.join(f'{l:.5f}' for l in ls), flush=True)
        if getattr(step, 'graph_error', None):
            print('graph_error:', step.graph_error)
    d = max(float((a - b).abs().max()) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    print('max |param difference| eager vs graph after 8 steps (different random draws):', d)


if __name__ == '__main__':
    try:
        main()
    except Exception:
        traceback.print_exc()
