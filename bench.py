#!/usr/bin/env python
"""bench.py - TSDF-volumes/sec of the GraspNeRF volumetric hot path on B200 (see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (oracle port)

A "step" = one pass of the hot path (NeuralRayRenderer.sample_volume given the encoders' feature maps: K1 -> K2a -> K2b)
over one scene of BASELINE.json configs[1] (6 views 288x512, 40^3 grid).  Steps cycle through a pool of 8 different
synthetic scenes (8 x 24.8 MB of inputs > the 126 MB L2), so no step finds its inputs in L2.
  value : volumes/s with inputs resident in HBM (device-timed, max over ranks, all ranks' volumes counted)
  e2e   : the same through graspnerf_b200.engine.VolumeEngine with pinned HOST inputs (H2D + kernels + D2H per step)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'TSDF-volumes/sec (6-view 288x512, 40^3 grid)'
V, H, W, R = 6, 288, 512, 40
POOL = 8
K1_BYTES_SURVEY = 4 * V * (3 * H * W + 2 * 32 * (H // 4) * (W // 4)) + 4 * R ** 3 * (V * 72 + 70)   # 153,284,608
K1_BYTES = 4 * V * (3 * H * W + 2 * 32 * (H // 4) * (W // 4)) + 4 * R ** 3 * (V * 72 + 2)           # what K1 moves: 135,876,608
K2_FLOPS = 2 * R ** 3 * (V * 28464 + 9104)                                                  # ~23.2 GFLOP


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback'}


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` captures
    (profiles/traffic.json, written by tools/ncu_summary.py); {} when no capture has been summarised yet."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    return json.load(open(p)) if os.path.exists(p) else {}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = [n for i, n in enumerate(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'])
                   if any(r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.rows[0][1]), 'reasons': reasons, 'samples': len(self.rows)}


def make_pool(n, seed0=0):
    from graspnerf_b200.synth import make_scene
    return [make_scene(seed=seed0 + s, num_views=V, h=H, w=W) for s in range(n)]


def cpu_reference_volumes(n_volumes, warmup=1, budget_s=40.0):
    """The reference algorithm for the path on the host cores: oracle/nr_oracle.sample_volume (a torch-CPU restatement
    of renderer.py:164-199, pinned to the real reference by tests/golden).  The reference itself is Python and cannot
    travel to the GPU box, so kind = "port".  This many-small-ops workload gets SLOWER with very many intra-op
    threads, so a short calibration picks the fastest of {8,16,32,64,all} threads ("all the host threads it can use")
    and the sample runs with that count, bounded to about `budget_s` seconds."""
    from oracle import nr_oracle as O
    from tests.helpers import golden_weights
    sd = golden_weights()
    sc = make_pool(1)[0]
    sct = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in sc.items()}
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], float('inf')
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            O.sample_volume(sd, sct)                      # warm-up at this thread count
            t0 = time.perf_counter()
            O.sample_volume(sd, sct)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
        torch.set_num_threads(best)
        for _ in range(warmup):
            O.sample_volume(sd, sct)
        n = max(1, min(n_volumes, int(budget_s / max(best_t, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(n):
            O.sample_volume(sd, sct)
        dt = time.perf_counter() - t0
    return n / dt, dt, n, best


def make_train_data(seed, dev):
    """One synthetic training sample of SURVEY.md 8d: configs[1]-sized scene + sdf_gt ~ U(-1,1), true_depth ~ U(0.2,0.8),
    512 query rays of view 0 with their colours (train_dataset.py:85), 64 random grasps."""
    from graspnerf_b200.synth import make_scene, make_query
    sc = make_scene(seed=seed, num_views=V, h=H, w=W)
    rng = np.random.default_rng(1000 + seed)
    ref = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in sc.items() if k not in ('img_feats', 'ray_feats')}
    ref['sdf_gt'] = torch.from_numpy(rng.uniform(-1, 1, (R, R, R)).astype(np.float32)).to(dev)
    ref['true_depth'] = torch.from_numpy(rng.uniform(0.2, 0.8, (V, 1, H, W)).astype(np.float32)).to(dev)
    que = {k: torch.from_numpy(v).to(dev) for k, v in make_query(sc, 512, seed).items() if isinstance(v, np.ndarray)}
    G = 64
    quat = rng.standard_normal((G, 2, 4)).astype(np.float32)
    quat /= np.linalg.norm(quat, axis=-1, keepdims=True)
    grasp = [torch.from_numpy(rng.integers(0, R, (G, 3))).to(dev), torch.from_numpy((rng.random(G) < 0.5).astype(np.float32)).to(dev),
             torch.from_numpy(quat).to(dev), torch.from_numpy(rng.uniform(0, 10, G).astype(np.float32)).to(dev)]
    return {'step': 0, 'ref_imgs_info': ref, 'que_imgs_info': que, 'src_imgs_info': ref, 'grasp_info': grasp}


def train_leg(args, dist, dev, world, rank, barrier, max_over_ranks):
    """configs[2]/[3]-style optimizer step (reported as an extra key, not the headline metric): `train_batch` scenes per
    GPU, GraspNeRF mirror forward with the SHIPPED configuration (render_rgb on: 512 rays x 40 coarse + 40 fine samples,
    40^3 volume, depth-mean head, cuDNN encoders, VGN), the four losses of nrvgn_sdf.yaml (render, depth, sdf + eikonal,
    vgn), backward through the hand-written reverse kernels, ONE all-reduce of the flat gradient bucket over the ranks, Adam."""
    from graspnerf_b200.network import name2network, NRVGN_SDF_CFG
    from graspnerf_b200.train import TrainStep
    cfg = dict(NRVGN_SDF_CFG)
    torch.manual_seed(0)
    net = name2network[cfg['network']](cfg).to(dev).train()
    step = TrainStep(net, lr=1e-4, dist=dist)
    nb = args.train_batch
    batch = [make_train_data(rank * nb + i, dev) for i in range(nb)]
    step(batch)                                               # warm-up: cuDNN autotune, allocator pools, kernel attributes
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    losses = [step(batch) for _ in range(args.train_steps)]
    ev1.record()
    barrier()
    ms = max_over_ranks([ev0.elapsed_time(ev1)], dist, dev)[0] / args.train_steps
    return {'value': world * nb / (ms / 1e3), 'unit': 'scenes/s', 'ms_per_step': ms, 'scenes_per_gpu': nb, 'global_batch': world * nb,
            'steps': args.train_steps, 'allreduce_bytes': step.bucket.nbytes if world > 1 else 0, 'loss': losses[-1],
            'what': 'GraspNeRF mirror fwd+bwd, shipped config (render_rgb on, 512 rays coarse+fine, 40^3 volume) + render/depth/sdf/eikonal/vgn losses + 1 NCCL all-reduce + Adam; 6x288x512'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    vps, dt, nvol, nthr = cpu_reference_volumes(args.steps, max(args.warmup, 1), budget_s=120.0)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': vps, 'unit': 'volumes/s', 'n_gpus': args.gpus, 'steps': nvol,
        'warmup': args.warmup, 'ms_per_step': 1e3 / vps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'configs[1]: 1 scene, 6x288x512, 40^3 grid, sample_volume given feature maps'},
        'cpu_baseline': {'value': vps, 'unit': 'volumes/s', 'cores': nthr, 'kind': 'port', 'host_cpus': os.cpu_count(),
                         'sample': f'{nvol} volumes of the workload in {dt:.1f} s, oracle/nr_oracle.sample_volume, torch CPU fp32, '
                                   f'{nthr} threads (fastest of a 8/16/32/64/all calibration)'},
        'e2e': {'value': vps, 'unit': 'volumes/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cpu-volumes', type=int, default=12, help='size of the bounded CPU-baseline sample')
    ap.add_argument('--train-batch', type=int, default=32, help='scenes per GPU of the extra training-step leg (0 = skip); 32 = BASELINE configs[2], and configs[3] (256 scenes) at --gpus 8')
    ap.add_argument('--train-steps', type=int, default=1)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs under ncu only)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    from graspnerf_b200 import ops
    from graspnerf_b200.engine import VolumeEngine, HostScene
    from tests.helpers import golden_weights

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: graspnerf_b200 has no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    W_ = max(args.warmup, 3)
    K = args.steps

    sd = golden_weights()              # random init of the reference architecture under torch.manual_seed(0)
    hw = ops.HeadWeights(sd, 'agg_net.', 'dist_decoder.', dev)
    from graspnerf_b200.shard import shard_scenes, max_over_ranks
    # every rank owns its own shard of the global scene stream (rank r takes scenes r, r+W, ...): no data-path collective
    pool = [make_pool(1, seed0=s)[0] for s in shard_scenes(POOL * world, rank, world)]
    scenes, bboxes, hosts = [], [], []
    for sc in pool:
        t = {k: torch.from_numpy(v).to(dev) for k, v in sc.items() if isinstance(v, np.ndarray)}
        s = ops.Scene(t['imgs'], t['img_feats'], t['ray_feats'], t['poses'], t['Ks'], t['depth_range'])
        scenes.append(s)
        bboxes.append(torch.tensor([sc['bbox3d'][0]], device=dev))
        hosts.append(HostScene(sc['imgs'], s.img_feats[0].cpu(), s.ray_feats[0].cpu(), sc['poses'], sc['Ks'], sc['depth_range'],
                               np.asarray(sc['bbox3d'][0], np.float32)))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    # one CUDA graph per pool scene (K1 -> K2a -> K2b captured once); a step = one graph launch
    graphs = [ops.VolumeGraph(scenes[i], hw, bboxes[i], R) for i in range(POOL)]
    for i in range(W_):
        graphs[i % POOL].replay()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        graphs[(W_ + i) % POOL].replay()
    ev1.record()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    # per-kernel durations.  CUDA event timestamps on this GPU tick at 4.096 us, so a single ~50 us launch cannot be timed
    # with its own event pair (round-1 numbers taken that way were quantised).  Each kernel is therefore launched 2*POOL
    # times back to back, each launch on ANOTHER scene's buffers (inputs 8 x 25 MB, records 8 x 110 MB: the working set
    # cycles past the 126 MB L2), between ONE event pair on the launching stream, behind a 1 GiB fill that flushes L2 and
    # lets the launches queue up.
    filler = torch.empty(1 << 28, device=dev, dtype=torch.float32)
    inter = []
    for i in range(POOL):
        s_, bb_ = scenes[i], bboxes[i]
        rec_, pt_ = ops.k1_forward(s_, hw, resolution=R, bbox_min=bb_)
        tok_ = ops.k2a_forward(rec_, pt_, hw, s_.depth_range, impl=ops.K2A_IMPL, want_pooled=False, want_tok=True, resolution=R, bbox_min=bb_)[3]
        inter.append((rec_, pt_, tok_))

    def k_time(fn, reps=3):
        out = []
        for _ in range(reps):
            for _f in range(6):           # ~1 ms of fills: the 16 launches below are queued before the GPU reaches them
                filler.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(2 * POOL):
                fn(i % POOL)
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / (2 * POOL))
        return float(np.median(out))
    kt = np.array([
        k_time(lambda i: ops.k1_forward(scenes[i], hw, resolution=R, bbox_min=bboxes[i])),
        k_time(lambda i: ops.k2a_forward(inter[i][0], inter[i][1], hw, scenes[i].depth_range, impl=ops.K2A_IMPL, want_pooled=False,
                                         want_tok=True, resolution=R, bbox_min=bboxes[i])),
        k_time(lambda i: ops.k2b_forward(None, hw, dn=R, resolution=R, bbox_min=bboxes[i], tok=inter[i][2]))])     # ms per launch: K1, K2a, K2b
    del filler, inter

    # ---------------- end-to-end timing (pinned host in, pinned host out) ----------------
    eng = VolumeEngine(hw, hosts[0], R, slots=3, device=dev)
    for i in range(W_):
        eng.submit(hosts[i % POOL])
    eng.drain()
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for i in range(K):
        _, fin = eng.submit(hosts[(W_ + i) % POOL], tag=i)
        if fin is not None:
            checksum += float(fin[1][0, 0, 0, 0, 0])
    for _, out in eng.drain():
        checksum += float(out[0, 0, 0, 0, 0])
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True              # sampled across both timed regions (device-resident and end-to-end)

    total_ms, e2e_ms = max_over_ranks([total_ms, e2e_ms], dist, dev)
    sampler.join(timeout=2)
    train = None
    h2d_bytes, d2h_bytes = eng.h2d_bytes, eng.d2h_bytes
    if args.train_batch > 0:
        del graphs, eng
        torch.cuda.empty_cache()
        try:
            train = train_leg(args, dist, dev, world, rank, barrier, max_over_ranks)
        except Exception as e:                                # the extra leg must never take the headline line down
            train = {'error': f'{type(e).__name__}: {e}'[:300]}
    if rank == 0:
        peaks = load_peaks()
        traffic = load_traffic()
        value = world * K / (total_ms / 1e3)
        k1_gbs = K1_BYTES / (kt[0] * 1e-3) / 1e9
        k2_tfs = K2_FLOPS / ((kt[1] + kt[2]) * 1e-3) / 1e12
        dominant_k2 = (kt[1] + kt[2]) >= kt[0]
        roof_k1 = {'kernel': 'gn_k1_kernel', 'bound': 'hbm', 'achieved': k1_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                   'frac': k1_gbs / peaks['hbm_gbs'], 'traffic': traffic.get('gn_k1_kernel'), 'us_per_launch': kt[0] * 1e3, 'peak_source': peaks['source'],
                   'algorithmic_bytes': K1_BYTES,
                   'achieved_survey_bytes': K1_BYTES_SURVEY / (kt[0] * 1e-3) / 1e9,
                   'note': 'achieved uses the bytes K1 itself must move (inputs once + 72-float record + 2 floats/point); '
                           'achieved_survey_bytes uses SURVEY 8d figure (153,284,608 B, counts mean/var that now live in K2a)'}
        roof_k2 = {'kernel': 'gn_k2a_tc3_kernel+gn_k2b_attn_kernel', 'bound': 'tensor', 'achieved': k2_tfs, 'peak': peaks['tf_sustained'],
                   'unit': 'TFLOP/s', 'frac': k2_tfs / peaks['tf_sustained'], 'traffic': traffic.get('gn_k2a_tc3_kernel'),
                   'us_per_launch': (kt[1] + kt[2]) * 1e3, 'peak_source': peaks['source'],
                   'note': 'algorithmic fp32 FLOPs of the reference semantics (SURVEY 8d: 23.2 GFLOP/volume) over the K2a+K2b time; '
                           'the kernel issues 3 fp16 MMAs per product (hi/lo split), so tensor-pipe activity is ~3x this fraction'}
        line = {
            'metric': METRIC, 'value': value, 'unit': 'volumes/s', 'n_gpus': world, 'steps': K, 'warmup': W_,
            'ms_per_step': total_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': 'configs[1]: 1 scene/step, 6x288x512, 40^3 grid, sample_volume given feature maps',
                       'l2': f'inputs cycle through {POOL} scenes x 24.8 MB (> 126 MB L2); no flush kernel in the timed region',
                       'parallelism': f'replicas x{world} (scenes sharded, no data-path collective)'},
            'roofline': roof_k2 if dominant_k2 else roof_k1,
            'roofline_k1': roof_k1, 'roofline_k2': roof_k2,
            'kernel_us': {'k1': kt[0] * 1e3, 'k2a': kt[1] * 1e3, 'k2b': kt[2] * 1e3},
            'e2e': {'value': world * K / (e2e_ms / 1e3), 'unit': 'volumes/s', 'h2d_bytes_per_step': h2d_bytes,
                    'd2h_bytes_per_step': d2h_bytes, 'api': 'graspnerf_b200.engine.VolumeEngine.submit (pinned host buffers)'},
            'gpu_launches': 3 * K, 'launch_mode': 'CUDA graph of the 3 kernels per scene (cudaGraphLaunch per step); kernel_us: each kernel launched 16x back to back over 8 scenes between one event pair (event clock ticks at 4.096 us)',
            'clocks': sampler.summary(),
            'checksum': checksum,
            'train_step': train,
        }
        if world == 1 and not args.no_cpu:
            vps, dt, nvol, nthr = cpu_reference_volumes(args.cpu_volumes, 1, budget_s=25.0)
            line['cpu_baseline'] = {'value': vps, 'unit': 'volumes/s', 'cores': nthr, 'kind': 'port', 'host_cpus': os.cpu_count(),
                                    'sample': f'{nvol} volumes of the same workload in {dt:.1f} s (oracle/nr_oracle.sample_volume, '
                                              f'torch CPU fp32, {nthr} threads = fastest of a 8/16/32/64/all calibration)'}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
