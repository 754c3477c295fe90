"""Where does one training step of the mirror go?  (development aid)  torch.profiler over 2 steps of bench.py's train leg."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from graspnerf_b200.network import name2network, NRVGN_SDF_CFG
from graspnerf_b200.train import TrainStep


def main():
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    net = name2network['grasp_nerf'](dict(NRVGN_SDF_CFG)).to(dev).train()
    step = TrainStep(net, lr=1e-4)
    batch = [bench.make_train_data(i, dev) for i in range(8)]
    step(batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(batch)
    torch.cuda.synchronize()
    print(f'wall per scene: {(time.perf_counter() - t0) / 8 * 1e3:.1f} ms')
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step(batch)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=60))
    print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=40, max_name_column_width=60))


if __name__ == '__main__':
    main()
