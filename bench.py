#!/usr/bin/env python
"""bench.py - TSDF-volumes/sec of the GraspNeRF volumetric hot path on B200 (see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's own implementation on the host cores

A "step" = one pass of the hot path (NeuralRayRenderer.sample_volume given the encoders' feature maps: K1 -> K2a -> K2b)
over one scene of BASELINE.json configs[1] (6 views 288x512, 40^3 grid).  Steps cycle through a pool of 8 different
synthetic scenes (8 x 28 MB of inputs > the 126 MB L2 once the 8 x 110 MB records are counted), so no step finds its inputs in L2.
  value : volumes/s with inputs resident in HBM (device-timed, max over ranks, all ranks' volumes counted)
  e2e   : the same through graspnerf_b200.engine.VolumeEngine with pinned HOST inputs (H2D + kernels + D2H per step)
Extra legs on the same line (not the headline): `full_forward` (the planner's whole GraspNeRF.forward, images in),
`highres` (BASELINE configs[4] shape), `train_step` (configs[2]/[3]).  Prints ONE JSON line on rank 0.

The CPU arm (`--impl reference`, and `cpu_baseline`) runs the UNMODIFIED reference from oracle/_ref (a verbatim copy made by
oracle/make_ref.py, kind "reference"); only if that copy is absent it falls back to the oracle port (kind "port").
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
WORKLOAD = 'configs[1]: 1 scene/step, 6x288x512, 40^3 grid, NeuralRayRenderer.sample_volume given the encoders\' feature maps'
V, H, W, R = 6, 288, 512, 40
POOL = 8


def k1_bytes(v, h, w, r, survey=False):
    """Algorithmic bytes of K1 per scene: inputs once + the 72-float record + 2 floats per point; survey=True
    is SURVEY 8d's figure, which also counts the 70 floats/point of mean/var that K2a produces here."""
    return 4 * v * (3 * h * w + 2 * 32 * (h // 4) * (w // 4)) + 4 * r ** 3 * (v * 72 + (70 if survey else 2))


def k2_flops(v, r):
    return 2 * r ** 3 * (v * 28464 + 9104)                        # SURVEY 8d, reference semantics


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


def make_pool(n, seed0=0, **kw):
    from graspnerf_b200.synth import make_scene
    return [make_scene(seed=seed0 + s, num_views=V, h=H, w=W, **kw) for s in range(n)]


def quantise_images(sc):
    """The planner's images are PNG bytes (main.py:166-171).  The synthetic U[0,1) images are quantised to uint8 once and
    BOTH arms see imgs = u8 / 255 in fp32 (color_map_forward), so the uint8 transport of the e2e leg changes no value."""
    u8 = np.clip(np.floor(sc['imgs'] * 256.0), 0, 255).astype(np.uint8)          # [V,3,H,W]
    sc = dict(sc)
    sc['imgs'] = u8.astype(np.float32) / np.float32(255.0)
    sc['imgs_u8'] = np.ascontiguousarray(u8.transpose(0, 2, 3, 1))               # [V,H,W,3] as imread gives them
    return sc


# ------------------------------------------------------------------------------------------------------------ CPU arm
def _pick_threads(fn):
    """This many-small-ops workload gets SLOWER with very many intra-op threads: a short calibration picks the fastest of
    {8,16,32,64,all} ("all the host threads it can use"); returns (threads, seconds per call at that count)."""
    ncpu = os.cpu_count() or 1
    best, best_t = None, float('inf')
    for c in sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu}):
        torch.set_num_threads(c)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best, best_t


def _timed(fn, n_max, per_call_s, budget_s, warmup=1):
    for _ in range(warmup):
        fn()
    n = max(1, min(n_max, int(budget_s / max(per_call_s, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return n, time.perf_counter() - t0


def _reference_net():
    """(net, kind): the UNMODIFIED reference's GraspNeRF under torch.manual_seed(0) (oracle/_ref or /root/reference), or
    (None, 'port') when no copy of the reference is available."""
    from oracle import ref_harness as RH
    if not RH.reference_available():
        return None, 'port'
    with RH.shims():
        _, net = RH.build_reference_net(0)
    return net, 'reference'


def cpu_hot_path(n_volumes, warmup=1, budget_s=40.0):
    """sample_volume (renderer.py:164-199) of the reference on the host cores, feature maps given."""
    from oracle import ref_harness as RH
    sc = quantise_images(make_pool(1)[0])
    sct = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in sc.items() if k != 'imgs_u8'}
    net, kind = _reference_net()
    with torch.no_grad():
        if net is not None:
            with RH.shims():
                fn = lambda: net.nr_net.sample_volume(dict(sct))
                nthr, per = _pick_threads(fn)
                n, dt = _timed(fn, n_volumes, per, budget_s, warmup)
            what = 'UNMODIFIED reference NeuralRayRenderer.sample_volume (oracle/_ref copy of src/nr)'
        else:
            from oracle import nr_oracle as O
            from graspnerf_b200.weights import seed0_weights
            sd = seed0_weights()
            fn = lambda: O.sample_volume(sd, sct)
            nthr, per = _pick_threads(fn)
            n, dt = _timed(fn, n_volumes, per, budget_s, warmup)
            what = 'oracle/nr_oracle.sample_volume (port: no copy of the reference on this box)'
    return {'value': n / dt, 'unit': 'volumes/s', 'cores': nthr, 'kind': kind, 'host_cpus': os.cpu_count(),
            'sample': f'{n} volumes of the same workload in {dt:.1f} s: {what}, torch CPU fp32, {nthr} threads '
                      f'(fastest of a 8/16/32/64/all calibration)'}


def _forward_data(sc, full=True):
    from graspnerf_b200.synth import make_query
    ref = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in sc.items() if k not in ('img_feats', 'ray_feats', 'imgs_u8')}
    q = {k: torch.from_numpy(v) for k, v in make_query(sc, 16, 7).items() if isinstance(v, np.ndarray)}
    return {'step': 0, 'eval': True, 'full_vol': True, 'ref_imgs_info': ref, 'que_imgs_info': q, 'src_imgs_info': ref}


def cpu_full_forward(budget_s=20.0):
    """GraspNeRF.forward, eval, render_rgb off (main.py:150,244-247) of the reference on the host cores: encoders +
    sample_volume + depth-mean head + VGN."""
    from oracle import ref_harness as RH
    net, kind = _reference_net()
    if net is None:
        return None
    sc = quantise_images(make_pool(1)[0])
    data = _forward_data(sc)
    net.nr_net.cfg['render_rgb'] = False
    with torch.no_grad(), RH.shims():
        fn = lambda: net(dict(data))
        nthr, per = _pick_threads(fn)
        n, dt = _timed(fn, 8, per, budget_s, 0)
    return {'value': n / dt, 'unit': 'volumes/s', 'cores': nthr, 'kind': kind, 'host_cpus': os.cpu_count(),
            'sample': f'{n} calls of the UNMODIFIED reference GraspNeRF.forward (eval, render_rgb off) in {dt:.1f} s, {nthr} threads'}


def cpu_train_scene(budget_s=40.0):
    """One training scene of the reference on the host cores: train-mode GraspNeRF.forward with the shipped configuration
    (512 rays coarse + fine, 40^3 volume, depth-mean head, VGN) + backward of a scalar of every output.  BASELINE.md 3:
    'time one reference fwd+bwd scene and extrapolate linearly to the batch' - the value is scenes/s of ONE scene."""
    from oracle import ref_harness as RH
    from graspnerf_b200.synth import make_query
    net, kind = _reference_net()
    if net is None:
        return None
    sc = make_pool(1)[0]
    ref = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in sc.items() if k not in ('img_feats', 'ray_feats')}
    rng = np.random.default_rng(1000)
    ref['true_depth'] = torch.from_numpy(rng.uniform(0.2, 0.8, (V, 1, H, W)).astype(np.float32))
    q = {k: torch.from_numpy(v) for k, v in make_query(sc, 512, 0).items() if isinstance(v, np.ndarray)}
    data = {'step': 0, 'full_vol': True, 'ref_imgs_info': ref, 'que_imgs_info': q, 'src_imgs_info': ref}
    net.nr_net.cfg['render_rgb'] = True
    net.train()
    torch.set_num_threads(min(os.cpu_count() or 1, 32))

    def one():
        net.zero_grad(set_to_none=True)
        out = net(dict(data))
        loss = out['volume'].square().mean() + out['pixel_colors_nr'].square().mean() + out['pixel_colors_nr_fine'].square().mean() \
            + 0.1 * (out['sdf_gradient_error'].mean() + out['sdf_gradient_error_fine'].mean()) + out['depth_mean'].mean() \
            + sum(t.square().mean() for t in out['vgn_pred'])
        loss.backward()
    with RH.shims():
        t0 = time.perf_counter()
        one()
        first = time.perf_counter() - t0
        n, dt = (1, first) if first > budget_s / 2 else _timed(one, 3, first, budget_s, 0)
    net.eval()
    return {'value': n / dt, 'unit': 'scenes/s', 'cores': torch.get_num_threads(), 'kind': kind, 'extrapolated': 'per-scene rate; a batch of B scenes is B sequential scenes in the reference (bs = 1, trainer.py:39)',
            'sample': f'{n} train-mode fwd+bwd scene(s) of the UNMODIFIED reference (512 rays coarse+fine + 40^3 volume + depth-mean + VGN) in {dt:.1f} s'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cb = cpu_hot_path(args.steps, max(args.warmup, 1), budget_s=120.0)
    vps = cb['value']
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': vps, 'unit': 'volumes/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 / vps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': WORKLOAD},
        'cpu_baseline': cb,
        'e2e': {'value': vps, 'unit': 'volumes/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    if not args.skip_full:
        ff = cpu_full_forward(30.0)
        if ff is not None:
            line['full_forward'] = {'value': ff['value'], 'unit': 'volumes/s', 'cpu_baseline': ff}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------ extra legs
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


def train_leg(args, dist, dev, world, rank, barrier, max_over_ranks, tf32=False):
    """configs[2]/[3]-style optimizer step (an extra key, not the headline metric): `train_batch` scenes per GPU, GraspNeRF
    mirror forward with the SHIPPED configuration (render_rgb on: 512 rays x 40 coarse + 40 fine samples, 40^3 volume,
    depth-mean head, cuDNN encoders, VGN), the four losses of nrvgn_sdf.yaml (render, depth, sdf + eikonal, vgn), backward
    through the hand-written reverse kernels, ONE all-reduce of the flat gradient bucket over the ranks, Adam.
    tf32: what cuDNN / cuBLAS may do in the torch-side encoders, VGN and per-ray head.  False (the reported value): plain fp32,
    like the CPU reference arm; True: PyTorch's default for convolutions on this GPU (what the reference itself would run
    with on an Ampere+ GPU), reported beside it as `tf32_convs`."""
    from graspnerf_b200.network import name2network, NRVGN_SDF_CFG
    from graspnerf_b200.train import TrainStep
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(NRVGN_SDF_CFG)
    torch.manual_seed(0)
    net = name2network[cfg['network']](cfg).to(dev).train()
    step = TrainStep(net, lr=1e-4, dist=dist, graph=True)
    nb = args.train_batch
    batch = [make_train_data(rank * nb + i, dev) for i in range(nb)]
    step(batch)                                               # warm-up: cuDNN autotune, allocator pools, kernel attributes
    step(batch)                                               # second step: captures the CUDA graph of one 8-scene group
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    losses = [step(batch) for _ in range(args.train_steps)]
    ev1.record()
    barrier()
    ms = max_over_ranks([ev0.elapsed_time(ev1)], dist, dev)[0] / args.train_steps
    return {'value': world * nb / (ms / 1e3), 'unit': 'scenes/s', 'ms_per_step': ms, 'scenes_per_gpu': nb, 'global_batch': world * nb,
            'steps': args.train_steps, 'allreduce_bytes': step.bucket.nbytes if world > 1 else 0, 'loss': losses[-1], 'cuda_graph': step._g is not None, 'cudnn_tf32': bool(tf32),
            'what': 'GraspNeRF mirror fwd+bwd, shipped config (render_rgb on, 512 rays coarse+fine, 40^3 volume) + render/depth/sdf/eikonal/vgn losses + 1 NCCL all-reduce + Adam; 6x288x512; forward+backward of each 8-scene group replayed as one CUDA graph'}


def highres_leg(args, dist, dev, world, rank, barrier, max_over_ranks, hw, peaks):
    """BASELINE configs[4] shape: 12 views 720x1280 (180x320 feature maps), 80^3 grid; `highres_scenes` scenes per GPU
    (configs[4] = 64 scenes over 8 GPUs = 8 per GPU), each scene one K1 -> K2a -> K2b pass (the reference cannot run this
    shape as shipped: hard-coded 40s).  Two different scenes alternate (2 x 310 MB of inputs + 2 x 1.8 GB records)."""
    from graspnerf_b200 import ops
    from graspnerf_b200.synth import make_scene
    v, h, w, r = 12, 720, 1280, 80
    scenes = []
    for s in range(2):
        sc = make_scene(seed=100 + rank * 2 + s, num_views=v, h=h, w=w, radius=0.55)
        t = {k: torch.from_numpy(x).to(dev) for k, x in sc.items() if isinstance(x, np.ndarray)}
        scenes.append((ops.Scene(t['imgs'], t['img_feats'], t['ray_feats'], t['poses'], t['Ks'], t['depth_range']),
                       torch.tensor([sc['bbox3d'][0]], device=dev)))
        del t
    n = args.highres_scenes
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    # per-kernel times of one eager pass (events between the launches), then the timed region: one CUDA graph per scene
    sc, bb = scenes[0]
    ops.sample_volume(sc, hw, bb, r)
    kt = np.full(3, np.inf)
    flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.float32)
    for _ in range(3):                                   # min of three eager passes (the first one still pays allocator / clock ramp)
        for _f in range(4):                              # ~1 ms of fills: L2 flushed, and the launches below queue up behind them
            flush.fill_(1.0)                             # (otherwise the event pairs also time the host's launch path)
        ev[0].record()
        rec, pt = ops.k1_forward(sc, hw, resolution=r, bbox_min=bb)
        ev[1].record()
        tok = ops.k2a_forward(rec, pt, hw, sc.depth_range, want_pooled=False, want_tok=True, resolution=r, bbox_min=bb)[3]
        ev[2].record()
        ops.k2b_forward(None, hw, dn=r, resolution=r, bbox_min=bb, tok=tok)
        ev[3].record()
        torch.cuda.synchronize()
        kt = np.minimum(kt, [ev[j].elapsed_time(ev[j + 1]) for j in range(3)])
    del rec, pt, tok, flush
    graphs = [ops.VolumeGraph(s_, hw, b_, r) for s_, b_ in scenes]
    for g in graphs:
        g.replay()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        graphs[i % 2].replay()
    e1.record()
    barrier()
    ms = max_over_ranks([e0.elapsed_time(e1)], dist, dev)[0]
    del graphs
    kb, kf = k1_bytes(v, h, w, r), k2_flops(v, r)
    return {'value': world * n / (ms / 1e3), 'unit': 'volumes/s', 'ms_per_volume': ms / n, 'scenes_per_gpu': n, 'global_batch': world * n,
            'workload': 'configs[4] shape: 12 views 720x1280, 80^3 grid (512 000 points, 6.1 M rows), sample_volume given feature maps',
            'kernel_ms': {'k1': kt[0], 'k2a': kt[1], 'k2b': kt[2]},
            'k1_gbs': kb / (kt[0] * 1e-3) / 1e9, 'k1_frac_of_hbm': kb / (kt[0] * 1e-3) / 1e9 / peaks['hbm_gbs'], 'k1_algorithmic_bytes': kb,
            'k2_tflops': kf / ((kt[1] + kt[2]) * 1e-3) / 1e12, 'k2_frac_of_bf16_sustained': kf / ((kt[1] + kt[2]) * 1e-3) / 1e12 / peaks['tf_sustained'],
            'k2_algorithmic_flops': kf}


def full_forward_leg(args, dist, dev, world, rank, barrier, max_over_ranks, pool_np):
    """The planner's whole network call (GraspNeRF.forward, eval, render_rgb off: main.py:150,244-247) through
    engine.ForwardEngine: uint8 images from pinned host memory -> 2-D encoders (K6 / K7, fp32-accurate) -> K1 -> K2a -> K2b ->
    depth-mean head -> VGN (K5) -> grasp post-processing on the device (K4) -> 7 volumes + grasp list back to pinned host
    memory.  Four scenes in flight, each slot's CUDA graph on a stream of its own (the scenes overlap on the GPU)."""
    from graspnerf_b200.engine import ForwardEngine, HostScene
    from graspnerf_b200.weights import seed0_model
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = seed0_model().to(dev).eval()
    net.nr_net.cfg['render_rgb'] = False
    hosts = [HostScene(sc['imgs_u8'], None, None, sc['poses'], sc['Ks'], sc['depth_range'], np.asarray(sc['bbox3d'][0], np.float32)) for sc in pool_np]
    post = dict(tsdf_thres_high=0.0, tsdf_thres_low=-0.85)                       # main.py:92-93
    eng = ForwardEngine(net, hosts[0], slots=4, device=dev, post_cfg=post)
    for i in range(8):
        eng.submit(hosts[i % len(hosts)])
    eng.drain()
    n = args.full_steps
    barrier()
    t0 = time.perf_counter()
    chk = 0.0
    for i in range(n):
        _, fin = eng.submit(hosts[i % len(hosts)], tag=i)
        if fin is not None:
            chk += float(fin[1][0][0, 0, 0, 0])
    for _, out in eng.drain():
        chk += float(out[0][0, 0, 0, 0])
    torch.cuda.synchronize()
    ms = max_over_ranks([(time.perf_counter() - t0) * 1e3], dist, dev)[0]
    return {'value': world * n / (ms / 1e3), 'unit': 'volumes/s', 'ms_per_volume': ms / n, 'steps': n,
            'h2d_bytes_per_step': eng.h2d_bytes, 'd2h_bytes_per_step': eng.d2h_bytes, 'cuda_graph': bool(eng.graphed), 'slots_in_flight': len(eng.slots), 'checksum': chk,
            'api': 'graspnerf_b200.engine.ForwardEngine.submit (pinned uint8 images in; tsdf/qual/rot/width volumes + grasp list out)',
            'what': 'GraspNeRF.forward eval, render_rgb off (main.py:150): 2-D encoders (tcgen05 convolutions K7 + fused norm/act/pad K6, two streams) + K1/K2a/K2b + depth-mean head (one launch) + VGN 3-D conv (K5) + process/select on the device (K4); one CUDA graph per slot, slots on concurrent streams'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cpu-volumes', type=int, default=12, help='size of the bounded CPU-baseline sample')
    ap.add_argument('--e2e-steps', type=int, default=400, help='submits of the end-to-end leg (at least this many regardless of --steps)')
    ap.add_argument('--train-batch', type=int, default=32, help='scenes per GPU of the extra training-step leg (0 = skip); 32 = BASELINE configs[2], and configs[3] (256 scenes) at --gpus 8')
    ap.add_argument('--train-steps', type=int, default=3)
    ap.add_argument('--highres-scenes', type=int, default=8, help='scenes per GPU of the configs[4]-shape leg (0 = skip); 8 = configs[4] at --gpus 8')
    ap.add_argument('--full-steps', type=int, default=200, help='submits of the full-forward leg')
    ap.add_argument('--skip-full', action='store_true')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline legs (profiling runs under ncu only)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    from graspnerf_b200 import ops
    from graspnerf_b200.engine import VolumeEngine, HostScene
    from graspnerf_b200.weights import seed0_weights

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: graspnerf_b200 has no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    from graspnerf_b200.engine import bind_to_gpu_numa
    numa_node = bind_to_gpu_numa(local) if world > 1 else None     # pinned staging buffers local to this rank's GPU
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    W_ = max(args.warmup, 3)
    K = args.steps

    sd = seed0_weights()               # random init of the reference architecture under torch.manual_seed(0) (bit-equal to the reference's)
    hw = ops.HeadWeights(sd, 'agg_net.', 'dist_decoder.', dev)
    from graspnerf_b200.shard import shard_scenes, max_over_ranks
    # every rank owns its own shard of the global scene stream (rank r takes scenes r, r+W, ...): no data-path collective
    pool = [quantise_images(make_pool(1, seed0=s)[0]) for s in shard_scenes(POOL * world, rank, world)]
    scenes, bboxes, hosts = [], [], []
    for sc in pool:
        t = {k: torch.from_numpy(v).to(dev) for k, v in sc.items() if isinstance(v, np.ndarray)}
        s = ops.Scene(t['imgs'], t['img_feats'], t['ray_feats'], t['poses'], t['Ks'], t['depth_range'])
        scenes.append(s)
        bboxes.append(torch.tensor([sc['bbox3d'][0]], device=dev))
        hosts.append(HostScene(sc['imgs_u8'], s.img_feats[0].cpu(), s.ray_feats[0].cpu(), sc['poses'], sc['Ks'], sc['depth_range'],
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
    # with its own event pair.  Each kernel is therefore launched 2*POOL times back to back, each launch on ANOTHER scene's
    # buffers (inputs 8 x 28 MB, records 8 x 110 MB: the working set cycles past the 126 MB L2), between ONE event pair on the
    # launching stream, behind ~1 ms of fills that flush L2 and let the launches queue up.
    filler = torch.empty(1 << 28, device=dev, dtype=torch.float32)
    inter = []
    for i in range(POOL):
        s_, bb_ = scenes[i], bboxes[i]
        rec_, pt_ = ops.k1_forward(s_, hw, resolution=R, bbox_min=bb_)
        tok_ = ops.k2a_forward(rec_, pt_, hw, s_.depth_range, want_pooled=False, want_tok=True, resolution=R, bbox_min=bb_)[3]
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
        k_time(lambda i: ops.k2a_forward(inter[i][0], inter[i][1], hw, scenes[i].depth_range, want_pooled=False,
                                         want_tok=True, resolution=R, bbox_min=bboxes[i])),
        k_time(lambda i: ops.k2b_forward(None, hw, dn=R, resolution=R, bbox_min=bboxes[i], tok=inter[i][2]))])     # ms per launch: K1, K2a, K2b
    del filler, inter

    # ---------------- end-to-end timing (pinned host in, pinned host out) ----------------
    KE = max(K, args.e2e_steps)
    eng = VolumeEngine(hw, hosts[0], R, slots=3, device=dev)
    for i in range(W_):
        eng.submit(hosts[i % POOL])
    eng.drain()
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for i in range(KE):
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
    h2d_bytes, d2h_bytes = eng.h2d_bytes, eng.d2h_bytes
    del graphs, eng
    torch.cuda.empty_cache()
    peaks = load_peaks()

    def guarded(fn, *a):                                      # an extra leg must never take the headline line down
        try:
            return fn(*a)
        except Exception as e:
            torch.cuda.synchronize()
            return {'error': f'{type(e).__name__}: {e}'[:300]}
    full = None if args.skip_full else guarded(full_forward_leg, args, dist, dev, world, rank, barrier, max_over_ranks, pool)
    torch.cuda.empty_cache()
    highres = guarded(highres_leg, args, dist, dev, world, rank, barrier, max_over_ranks, hw, peaks) if args.highres_scenes > 0 else None
    torch.cuda.empty_cache()
    train = guarded(train_leg, args, dist, dev, world, rank, barrier, max_over_ranks) if args.train_batch > 0 else None
    if train is not None and 'error' not in train:
        torch.cuda.empty_cache()
        t32 = guarded(train_leg, args, dist, dev, world, rank, barrier, max_over_ranks, True)
        train['tf32_convs'] = {k: t32.get(k) for k in ('value', 'ms_per_step', 'loss', 'error') if k in t32}
        torch.backends.cudnn.allow_tf32 = False
    if rank == 0:
        traffic = load_traffic()
        value = world * K / (total_ms / 1e3)
        kb = k1_bytes(V, H, W, R)
        k1_gbs = kb / (kt[0] * 1e-3) / 1e9
        k2_tfs = k2_flops(V, R) / ((kt[1] + kt[2]) * 1e-3) / 1e12
        dominant_k2 = (kt[1] + kt[2]) >= kt[0]
        roof_k1 = {'kernel': 'gn_k1_kernel', 'bound': 'hbm', 'achieved': k1_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                   'frac': k1_gbs / peaks['hbm_gbs'], 'traffic': traffic.get('gn_k1_kernel'), 'us_per_launch': kt[0] * 1e3, 'peak_source': peaks['source'],
                   'algorithmic_bytes': kb,
                   'achieved_survey_bytes': k1_bytes(V, H, W, R, survey=True) / (kt[0] * 1e-3) / 1e9,
                   'note': 'achieved uses the bytes K1 itself must move (inputs once + 72-float record + 2 floats/point); '
                           'achieved_survey_bytes uses SURVEY 8d figure (153,284,608 B, counts mean/var that now live in K2a)'}
        # the kernels are timed in isolation (16 back-to-back launches) -> the BURST bf16 peak is the denominator
        roof_k2 = {'kernel': 'gn_k2a_tc3_kernel+gn_k2b_attn_kernel', 'bound': 'tensor', 'achieved': k2_tfs, 'peak': peaks['tf_burst'],
                   'unit': 'TFLOP/s', 'frac': k2_tfs / peaks['tf_burst'], 'traffic': traffic.get('gn_k2a_tc3_kernel'),
                   'us_per_launch': (kt[1] + kt[2]) * 1e3, 'peak_source': peaks['source'], 'peak_kind': 'burst (kernel timed alone)',
                   'note': 'algorithmic fp32 FLOPs of the reference semantics (SURVEY 8d: 23.2 GFLOP/volume) over the K2a+K2b time; '
                           'the kernel issues 3 fp16 MMAs per product (hi/lo split), so tensor-pipe activity is ~3x this fraction'}
        line = {
            'metric': METRIC, 'value': value, 'unit': 'volumes/s', 'n_gpus': world, 'steps': K, 'warmup': W_,
            'ms_per_step': total_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD,
                       'l2': f'inputs cycle through {POOL} scenes x (28 MB inputs + 110 MB record) (> 126 MB L2); no flush kernel in the timed region',
                       'parallelism': f'replicas x{world} (scenes sharded, no data-path collective)' + (f', rank 0 bound to NUMA node {numa_node}' if numa_node is not None else ''),
                       'images': 'U[0,1) synthetic images quantised to uint8 once; both arms compute on u8/255 (main.py:170)'},
            'roofline': roof_k2 if dominant_k2 else roof_k1,
            'roofline_k1': roof_k1, 'roofline_k2': roof_k2,
            'kernel_us': {'k1': kt[0] * 1e3, 'k2a': kt[1] * 1e3, 'k2b': kt[2] * 1e3},
            'e2e': {'value': world * KE / (e2e_ms / 1e3), 'unit': 'volumes/s', 'h2d_bytes_per_step': h2d_bytes,
                    'd2h_bytes_per_step': d2h_bytes, 'steps': KE,
                    'api': 'graspnerf_b200.engine.VolumeEngine.submit (pinned host buffers: uint8 RGB images + fp32 fused feature maps in, fp32 volume out)'},
            'gpu_launches': 3 * K, 'launch_mode': 'CUDA graph of the 3 kernels per scene (cudaGraphLaunch per step); kernel_us: each kernel launched 16x back to back over 8 scenes between one event pair (event clock ticks at 4.096 us)',
            'clocks': sampler.summary(),
            'checksum': checksum,
            'full_forward': full, 'highres': highres, 'train_step': train,
        }
        if world == 1 and not args.no_cpu:
            line['cpu_baseline'] = cpu_hot_path(args.cpu_volumes, 1, budget_s=25.0)
            if isinstance(full, dict) and 'error' not in full:
                full['cpu_baseline'] = guarded(cpu_full_forward, 15.0)
            if isinstance(train, dict) and 'error' not in train:
                train['cpu_baseline'] = guarded(cpu_train_scene, 30.0)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
