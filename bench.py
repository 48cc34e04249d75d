#!/usr/bin/env python
"""Benchmark of the fused PINN training step (BASELINE.json metric: collocation points / second
through forward + NS-Darcy residual + backward), one process per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm (CPU oracle port) on the host cores

A step = one pass of the hot path over one batch of synthetic geometries: zero gradients, encode
(FPS / ball query / set abstraction), jet forward on internal + boundary points, fused residual,
reverse pass to every parameter gradient, gradient all-reduce (N > 1) and the fused Adam update.
Headline workload at every N: BASELINE config 2 -- PIPN++ (examples/abc, 3-D), 1500 / 1000 / 700 points,
32 geometries PER GPU (weak scaling), dropout on, laplacian='reference' (the operator the
reference's training_step computes as written).

`value`  : whole-job collocation points/s with the batch already resident in HBM.
`e2e`    : the same metric through the public API (model.training_step(batch) + loss.backward()
           + optimizer.step()) with the batch copied from pinned host memory and the loss read
           back every step.
`roofline`: the dominant kernel family of the step, timed per launch with CUDA events.
`cpu_baseline`: the oracle (a port of the reference's algorithm: D + D*D + 1 reverse sweeps and a
           double backward in torch CPU) on a bounded sample of the same workload.
`configs`: the other BASELINE.json configs at this N (config 3 and 5 strong-scaled: 64 / 256 geometries in total split
           over the ranks; config 4 swept over geometries per GPU), device-resident, each with its dominant-family
           roofline.  `--no-configs` skips them.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import pcfd_import  # noqa: E402

pcfd_import.load()
from porous_cfd_b200 import synthetic  # noqa: E402

N_BATCHES = 4            # distinct synthetic batches cycled through
METRIC = 'PINN train collocation points/sec (fwd+NS-Darcy residual+bwd)'

# BASELINE.json configs 1-5 (config 2 is the headline: the one the metric is quoted on for one GPU).  The others are
# selectable with --config and reported in the `configs` block of the default run; their parity is covered by tests/.
WORKLOADS = {
    'abc_pipn': dict(shape=dict(n_internal=1500, n_boundary=1000, n_obs=700), batch=13,
                     what='PIPN abc 3-D, vanilla (examples/abc/train.py:26-34; max-pool coupling terms included)'),
    'abc_pipn_pp': dict(shape=dict(n_internal=1500, n_boundary=1000, n_obs=700), batch=32,
                        what='PIPN++ abc 3-D (examples/abc/train.py:36-49)'),
    'duct_pigano': dict(shape=dict(n_internal=1500, n_boundary=1000, n_obs=700), batch=64,
                        what='PI-GANO duct_variable_boundary 2-D (examples/duct_variable_boundary/train.py:28-37)'),
    'windbreaks_pigano_pp': dict(shape=dict(n_internal=16384, n_boundary=8192, n_obs=4096), batch=2,
                                 what='PI-GANO++ windbreaks 3-D (examples/windbreaks/train.py:38-52)'),
    'manufactured_pipn_pp': dict(shape=dict(n_internal=4096, n_boundary=1024, n_obs=0), batch=32,
                                 what='PIPN++ manufactured solutions 2-D (examples/manufactured_solutions/train.py:19-27)'),
}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'tflops_burst': p['bf16_tflops'], 'tflops_sustained': p['bf16_tflops_sustained'],
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i',
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': mhz[len(mhz) // 2] if mhz else None,
                'sm_max_mhz': int(self.samples[0][1]) if self.samples[0][1].isdigit() else None, 'reasons': reasons}


def make_model(config: str, device, train=True):
    from porous_cfd_b200 import factory
    spec = synthetic.model_spec(config)
    torch.manual_seed(3)
    model = factory.build_model(spec).to(device)
    return (model.train() if train else model.eval()), spec


def shape_text(shape) -> str:
    return f"{shape['n_internal']}/{shape['n_boundary']}/{shape['n_obs']}"


# ------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------

def cpu_reference_step_rate(config: str, shape: dict, steps: int, warmup: int, n_geom: int):
    """Times the oracle's training_step + backward (CPU, all host threads) on `n_geom` geometries of a workload.
    Returns (points/s, ms/step, threads)."""
    from oracle import pinn_oracle
    from porous_cfd_b200 import factory
    spec = synthetic.model_spec(config)
    torch.manual_seed(3)
    params = {k: v.detach().clone() for k, v in factory.build_model(spec).state_dict().items()}
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    data, labels, domain = synthetic.make_batch(spec['layout'], n_geom, seed=8421, **shape)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        pinn_oracle.step_with_grads(spec, params, data, labels, domain, 'reference', training=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return n_geom * shape['n_internal'] / (ms / 1e3), ms, threads


def run_reference(args, config: str, shape: dict, batch: int):
    """--impl reference: the reference's CPU algorithm on the host cores, on THIS arm's config: every step processes all
    `batch` geometries of one GPU's share of the workload (same config as the GPU arm; the step count is bounded because a
    step takes seconds and is echoed in the line).  kind 'port': the reference itself needs lightning / torch_geometric /
    torch_cluster, which this image does not have and which cannot travel to the GPU box; oracle/pinn_oracle.py is pinned
    bit for bit on the unmodified reference by tests/golden/make_golden.py."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), 1
    pts, ms, threads = cpu_reference_step_rate(config, shape, steps, warmup, batch)
    extra = []
    if config == 'abc_pipn_pp' and not args.no_configs:
        # the reference's own CPU-runnable case (config 1, vanilla PIPN) at B = 1 and B = 13
        for b, st in ((1, 3), (13, 2)):
            p1, m1, _ = cpu_reference_step_rate('abc_pipn', WORKLOADS['abc_pipn']['shape'], st, 1, b)
            extra.append({'config': 1, 'workload': f"abc_pipn: {WORKLOADS['abc_pipn']['what']}, "
                                                   f"{shape_text(WORKLOADS['abc_pipn']['shape'])} points, {b} geometries per step",
                          'value': p1, 'unit': 'points/s', 'ms_per_step': m1, 'steps': st, 'warmup': 1, 'cores': threads,
                          'kind': 'port'})
    line = {'impl': 'reference', 'metric': METRIC,
            'value': pts, 'unit': 'points/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': ms,
            'steps_requested': args.steps, 'warmup_requested': args.warmup,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{config}: {WORKLOADS[config]["what"]}, {shape_text(shape)} points, '
                                   f'{batch} geometries per GPU', 'global_batch': batch,
                       'laplacian': 'reference', 'dropout': 'on',
                       'note': 'one host, all cores: the CPU arm runs one GPU\'s share of the batch whatever --gpus says'},
            'cpu_baseline': {'value': pts, 'unit': 'points/s', 'cores': threads, 'kind': 'port',
                             'sample': f'{batch} geometries x {shape["n_internal"]} collocation points per step '
                                       f'(the full per-GPU batch), {steps} timed steps after {warmup} warm-up'},
            'e2e': {'value': pts, 'unit': 'points/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'configs': extra}
    _emit(line)


_RESULT_FD = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner with
    NCCL_DEBUG=VERSION, which this image sets), so file descriptor 1 is pointed at stderr for the whole process and the
    result line goes to a duplicate of the original stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (json.dumps(line) + '\n').encode())


def geometry_cache_leg(model, trainer, dev_batches, labels, b: int, ni: int, steps: int) -> dict:
    """SURVEY 8f rank 4, second half, NEVER part of the headline: training steps fed from an HBM-resident dataset
    (DeviceFoamDataset.batch -> model.training_step [CUDA graph] -> backward -> Adam), once with FPS / ball query inside
    every step and once with the per-geometry cache (valid because the FPS start is fixed, see build_geometry_cache)."""
    from porous_cfd_b200.dataset.device_dataset import DeviceFoamDataset
    data = torch.cat([d.data for d in dev_batches])
    domain = {k: torch.cat([d.domain[k] for d in dev_batches]) for k in dev_batches[0].domain}
    gen = torch.Generator().manual_seed(1)
    ids = [torch.randperm(data.shape[0], generator=gen)[:b].cuda() for _ in range(8)]
    model.cuda_graph = True
    model.pipeline_geometry = False
    params = list(model.parameters())
    out = {}
    for mode in ('uncached', 'cached'):
        ds = DeviceFoamDataset(data, labels, domain)
        if mode == 'cached':
            t0 = time.perf_counter()
            ds.build_geometry_cache(model)
            torch.cuda.synchronize()
            out['build_ms'] = 1e3 * (time.perf_counter() - t0)
            out['cached_geometries'] = len(ds)

        def one(i):
            loss = model.training_step(ds.batch(ids[i % 8]), i)
            for p in params:
                p.grad = None
            loss.backward()
            trainer.step(model.executor.last_flat_grad)

        for i in range(4):
            one(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            one(4 + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[mode] = {'ms_per_step': ms, 'points_per_s': b * ni / (ms / 1e3)}
    out['note'] = ('both legs: batches collated on the device from an HBM-resident dataset, step replayed from a CUDA graph, no '
                   'one-step-ahead geometry pipeline; cached = FPS / ball query taken from the per-geometry cache (exact only '
                   'with the deterministic FPS start this build uses; the reference re-draws a random start every step)')
    return out


def train_loop_leg(config: str, dev_batches, labels, b: int, ni: int, tmp_dir: str) -> dict:
    """The reference-facing training loop itself (common.training.train: DataLoader + collate_fn on the host, pinned
    upload one batch ahead, step replayed from a CUDA graph, Adam), on an in-memory dataset of this workload's geometries:
    wall-clock points/s of whole epochs, host side included (reference loop: common/training.py:50-85)."""
    from argparse import Namespace
    from porous_cfd_b200.common.training import train
    from porous_cfd_b200.dataset.foam_data import FoamData

    class Mem(torch.utils.data.Dataset):
        def __init__(self, items):
            self.items = items

        def __len__(self):
            return len(self.items)

        def __getitem__(self, i):
            return self.items[i]

    items = []
    for d in dev_batches:
        data = d.data.cpu()
        dom = {k: v.cpu() for k, v in d.domain.items()}
        items += [FoamData(data[i], labels, {k: v[i] for k, v in dom.items()}) for i in range(data.shape[0])]
    steps = (len(items) + b - 1) // b
    out = {'epochs': 6, 'steps_per_epoch': steps, 'geometries': len(items),
           'note': 'train(): CUDA-graph step + Adam; epochs 3-6 of 6 (the first two run eagerly / capture).  host_loader: host '
                   'DataLoader (num_workers = 0) + collate_fn, pinned upload one batch ahead; device_dataset: the training set '
                   'resident in HBM, collation on the device (args.device_dataset = True)'}
    for mode, extra in (('host_loader', {}), ('device_dataset', {'device_dataset': True})):
        model, _ = make_model(config, torch.device('cuda', torch.cuda.current_device()))
        secs = []
        args = Namespace(n_internal=ni, n_boundary=0, n_observations=0, batch_size=b, precision='32', epochs=6,
                         logs_dir=os.path.join(tmp_dir, mode), train_dir='', val_dir='', model=config, name='bench', checkpoint=None,
                         loss_scaler='fixed', epoch_seconds=secs, **extra)
        try:
            train(args, model, Mem(items), Mem([]))
            steady = sorted(secs[2:])
            sec = steady[len(steady) // 2]
            out[mode] = {'median_epoch_ms': 1e3 * sec, 'ms_per_step': 1e3 * sec / steps, 'points_per_s': len(items) * ni / sec}
        except Exception as e:
            import traceback
            out[mode] = {'error': f'{type(e).__name__}: {str(e)[:200]}', 'trace': traceback.format_exc()[-1500:]}
            torch.cuda.synchronize()
    return out


def ingest_leg(dev_batches, labels, n_internal, dims, peaks):
    """SURVEY 8f rank 4, outside the timed step: collation of HBM-resident geometries (pcfd_gather_blocks) and the
    signed-distance feature (pcfd_sdf_feature) on this workload's shapes, CUDA events on the launching stream."""
    from porous_cfd_b200.dataset.device_dataset import DeviceFoamDataset
    b, n, f = dev_batches[0].data.shape
    copies = max(1, int(160e6 // (len(dev_batches) * b * n * f * 4)) + 1)        # resident set larger than the 126 MB L2
    data = torch.cat([d.data for d in dev_batches] * copies)
    domain = {k: torch.cat([d.domain[k] for d in dev_batches] * copies) for k in dev_batches[0].domain}
    ds = DeviceFoamDataset(data, labels, domain)
    gen = torch.Generator().manual_seed(0)
    ids = [torch.randperm(len(ds), generator=gen)[:b].cuda() for _ in range(8)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(3):
        ds.batch(ids[i])
    ev[0].record()
    for i in range(24):
        ds.batch(ids[i % 8])
    ev[1].record()
    torch.cuda.synchronize()
    ms_collate = ev[0].elapsed_time(ev[1]) / 24
    nbytes = 2.0 * (b * n * f * 4 + sum(v.shape[1] for v in domain.values()) * b * 8)
    sub = DeviceFoamDataset(dev_batches[0].data.clone(), labels, dev_batches[0].domain)
    for _ in range(2):
        sub.add_sdf()
    ev[0].record()
    for _ in range(5):
        sub.add_sdf()
    ev[1].record()
    torch.cuda.synchronize()
    ms_sdf = ev[0].elapsed_time(ev[1]) / 5
    pairs = float(b) * n * (n - n_internal)
    return {'resident_geometries': len(ds), 'resident_mb': data.numel() * 4 / 1e6,
            'collate': {'ms_per_batch': ms_collate, 'gbs': nbytes / (ms_collate / 1e3) / 1e9,
                        'frac_of_hbm_peak': nbytes / (ms_collate / 1e3) / 1e9 / peaks['hbm_gbs'],
                        'bytes': 'read + write of the batch tensor and every sub-domain id tensor',
                        'note': 'one launch per batch; a batch is a few MB, so this loop is bounded by the host-side '
                                'dispatch (output allocations + one ctypes call), not by HBM; it replaces the '
                                'host-to-device copy of the same bytes'},
            'sdf': {'ms_per_batch': ms_sdf, 'geometries': b, 'distance_pairs_per_s': pairs / (ms_sdf / 1e3),
                    'bound': 'fp32 FMA (n x n_boundary distance evaluations per geometry)'}}


# ------------------------------------------------------------------------------------------------
# one workload on this rank's GPU
# ------------------------------------------------------------------------------------------------

class Env:
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        self.device = torch.device('cuda', self.local)
        torch.cuda.set_device(self.device)
        if self.world > 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=self.device)
        self.args = args
        self.peaks = measured_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return vals
        t = torch.tensor(list(vals), device=self.device, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return tuple(float(v) for v in t)


def family_roofline(fam: dict, prof_steps: int, peaks: dict, engine_name: str) -> dict:
    """The `roofline` object of a workload from the per-family CUDA-event timings of the profiled leg."""
    ms_prof = sum(v['ms'] for v in fam.values())
    gemm = {k: v for k, v in fam.items() if k.startswith('jet_')}
    top_name = max(gemm, key=lambda k: gemm[k]['ms'])
    top = gemm[top_name]
    # The jet layers of these workloads are narrow (k, n <= 512): their arithmetic intensity, k*n / (2*(k+n)) FLOP per
    # byte = 27 for the widest layer of config 2, is below the 3xTF32 ridge of the B200 (~370 TFLOP/s effective /
    # 6.5 TB/s = 57), so the bounding roofline of the dominant family is HBM bandwidth; the tensor-pipe figure is beside it.
    achieved_gbs = top['bytes'] / (top['ms'] / 1e3) / 1e9
    achieved_tf = top['work'] / (top['ms'] / 1e3) / 1e12
    all_flops = sum(v['work'] for v in gemm.values())
    all_bytes = sum(v['bytes'] for v in gemm.values())
    all_ms = sum(v['ms'] for v in gemm.values())
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')      # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(top_name)
    roofline = {'bound': 'hbm', 'kernel': top_name, 'achieved': achieved_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': achieved_gbs / peaks['hbm_gbs'], 'traffic': traffic,
                'peak_source': peaks['source'] + ', copy bandwidth',
                'algorithmic_bytes_per_launch': top['bytes'] / top['launches'],
                'launches': top['launches'], 'avg_launch_ms': top['ms'] / top['launches'],
                'share_of_step': top['ms'] / ms_prof, 'engine': engine_name,
                'timing': 'CUDA events around every C-ABI call of an eagerly launched, single-stream step whose launches are all '
                          'queued behind a spin kernel first (no host-side gaps between the kernels); the captured step is timed in `value`',
                'tensor': {'achieved_tflops': achieved_tf, 'mma_tflops_3xtf32': 3 * achieved_tf,
                           'peak_bf16_tflops': peaks['tflops_sustained'], 'frac_of_bf16_peak': 3 * achieved_tf / peaks['tflops_sustained']},
                'all_jet_gemms': {'tflops': all_flops / (all_ms / 1e3) / 1e12,
                                  'gbs': all_bytes / (all_ms / 1e3) / 1e9,
                                  'hbm_frac': all_bytes / (all_ms / 1e3) / 1e9 / peaks['hbm_gbs'],
                                  'share_of_step': all_ms / ms_prof},
                'families_ms_per_step': {k: round(v['ms'] / prof_steps, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]['ms'])},
                'families_hbm_frac': {k: round(v['bytes'] / (v['ms'] / 1e3) / 1e9 / peaks['hbm_gbs'], 4) for k, v in gemm.items()}}
    hbm = {}
    for k in ('residual_loss', 'segmax_fwd', 'segmax_bwd', 'ball_query', 'sa_gather'):
        if k in fam:
            hbm[k] = {'gbs': fam[k]['work'] / (fam[k]['ms'] / 1e3) / 1e9, 'frac': fam[k]['work'] / (fam[k]['ms'] / 1e3) / 1e9 / peaks['hbm_gbs']}
    roofline['hbm_kernels'] = hbm
    if 'fps' in fam:
        roofline['fps'] = {'ms_per_launch': fam['fps']['ms'] / fam['fps']['launches'], 'note': 'latency-bound (sequential sampling)'}
    return roofline


def run_workload(env: Env, config: str, shape: dict, b_per_gpu: int, steps: int, warmup: int, micro: int = 1,
                 e2e: bool = True, seed_base: int = 0) -> dict:
    """Times `steps` optimizer steps of one workload.  A step processes `b_per_gpu` geometries on every rank, as `micro`
    micro-batches of b_per_gpu / micro geometries whose gradients accumulate (one all-reduce + one Adam update per step)."""
    from porous_cfd_b200 import _lib, ops
    from porous_cfd_b200.common.training import FlatAdamTrainer
    from porous_cfd_b200.dataset.foam_data import FoamData
    args, world, rank, device = env.args, env.world, env.rank, env.device
    dist = env.dist
    assert b_per_gpu % micro == 0
    b_micro = b_per_gpu // micro
    W, K = max(3, warmup), max(1, steps)
    engine = 0 if ops.FORCE_FFMA else 2
    engine_name = {0: 'fp32 FFMA (reference engine)', 2: 'TMA + tcgen05 3xTF32, warp-specialised'}[engine]
    ops.AUDIT = True          # wide layers must run on the tensor cores: fallbacks are recorded and reported
    ops.FALLBACKS.clear()

    model, spec = make_model(config, device)
    trainer = FlatAdamTrainer(model)
    trainer.accumulate = micro
    ex = model.executor
    ni = shape['n_internal']

    # distinct batches per rank, pinned on the host and resident on the device
    host, dev_batches = [], []
    n_batches = N_BATCHES if b_micro * (ni + shape['n_boundary']) < 4_000_000 else 2
    for i in range(n_batches):
        data, labels, domain = synthetic.make_batch(spec['layout'], b_micro, seed=seed_base + 1000 * rank + i, **shape)
        hb = FoamData(data, labels, domain)
        if e2e:
            hb = hb.pin_memory()
            host.append(hb)
        dev_batches.append(hb.to(device))
    h2d_bytes = dev_batches[0].data.numel() * 4 + sum(v.numel() * 8 for v in dev_batches[0].domain.values())

    # ---- (1) device-resident throughput: CUDA graph of the whole step when possible ---------
    static = FoamData(torch.empty_like(dev_batches[0].data), labels,
                      {k: torch.empty_like(v) for k, v in dev_batches[0].domain.items()})
    pipelined = (not args.no_pipeline) and (not args.no_graph) and ex.uses_geometry()
    geo = None

    def load_static(src: FoamData, dst: FoamData = static):
        names = list(src.domain)      # the batch tensor and every sub-domain's ids in one launch
        ops.copy_blocks_multi([src.data] + [src.domain[k] for k in names], [dst.data] + [dst.domain[k] for k in names])

    def finish_step():
        trainer.reduce_gradients()
        trainer.step()

    load_static(dev_batches[0])
    l0 = _lib.launches
    model.fused_step(static, args.laplacian)          # sizes the workspaces before any capture
    launches_per_micro = _lib.launches - l0
    finish_step()
    torch.cuda.synchronize()
    graph = None
    tail_in_graph = False
    if not args.no_graph:
        # the optimizer tail (all-reduce at N > 1, Adam) is part of the graph when the step is one micro-batch; NCCL
        # collectives are capturable, if this build refuses the graph is re-captured without the tail
        for with_tail in ((True, False) if micro == 1 else (False,)):
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):
                        model.fused_step(static, args.laplacian, accumulate=micro > 1)
                        if micro == 1:
                            finish_step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                if pipelined:
                    # geometry of the NEXT batch (positions only: FPS, ball query, edge slots) as an independent branch of
                    # the graph; this batch's geometry sits in static buffers filled by the previous replay
                    geo = ex.geometry(static.data, labels, static.domain)
                    pos_next = ex.geometry_positions(dev_batches[1 % n_batches].data, labels, dev_batches[1 % n_batches].domain)
                    geo_side = torch.cuda.Stream()
                    torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                gc.collect()          # no dead model (and its CUDA graph / memory pool) may be collected inside the capture
                gc.disable()
                with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                    if pipelined:
                        cap = torch.cuda.current_stream()
                        geo_side.wait_stream(cap)
                        with torch.cuda.stream(geo_side):
                            geo_next = ex.geometry(None, labels, None, pos=pos_next)
                    graph_res = model.fused_step(static, args.laplacian, geo=geo, accumulate=micro > 1)
                    if with_tail:
                        finish_step()      # all-reduce (N > 1) + fused Adam on the flat buffers, inside the graph
                    if pipelined:
                        cap.wait_stream(geo_side)
                        for cur_l, new_l in zip(geo, geo_next):
                            for gk in cur_l:
                                cur_l[gk].copy_(new_l[gk])
                tail_in_graph = with_tail
                gc.enable()
                break
            except Exception as e:  # report and fall back (still the CUDA path)
                gc.enable()
                if rank == 0:
                    print(f'[bench] CUDA graph capture failed (tail in graph: {with_tail}; {type(e).__name__}: {e})', file=sys.stderr)
                graph = None
                torch.cuda.synchronize()

    def micro_step(j):
        if graph is not None:
            load_static(dev_batches[j % n_batches])
            if pipelined:      # the geometry branch reads only the sampled positions of the next batch: one gather launch
                nb_ = dev_batches[(j + 1) % n_batches]
                ex.geometry_positions(nb_.data, labels, nb_.domain, out=pos_next)
            graph.replay()
            return graph_res
        return model.fused_step(dev_batches[j % n_batches], args.laplacian, accumulate=micro > 1)

    def step(i):
        if micro > 1:
            ops.zero_(ex.flat_grad)
        for m in range(micro):
            res = micro_step(i * micro + m)
        if not tail_in_graph:
            finish_step()
        return res

    for i in range(W):
        step(i)
    env.barrier()
    sampler = ClockSampler(env.local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        last = step(W + i)
    e1.record()
    env.barrier()
    ms_total = e0.elapsed_time(e1)
    loss_value = float(last.loss)

    # ---- (2) end to end through the public API, host batches, loss read back ----------------
    ms_e2e = None
    if e2e:
        params = list(model.parameters())
        for p in params:
            p.grad = None
        # The call a user makes: model.training_step(batch) + loss.backward() + optimizer.step() (+ float(loss)), with
        # model.cuda_graph = True (the fused step replays from a CUDA graph) and the NEXT step's pinned-host -> device
        # copy issued on a copy stream while this step computes (every step still performs one full H2D copy of a batch
        # inside the timed region, and reads its own loss back).
        model.cuda_graph = True
        copy_stream = torch.cuda.Stream()

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                b = model.transfer_batch_to_device(host[i % n_batches], device)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return b, ev

        model.pipeline_geometry = pipelined
        # with the geometry pipeline the graph of step i also reads the positions of batch i+1, so the copy of batch i+1
        # was issued during step i-1 and this step issues the copy of batch i+2: still one batch copied per step
        depth = 2 if pipelined else 1
        pending = {'queue': [prefetch(j) for j in range(depth)]}
        loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        loss_ev = [None, None]

        def read_loss(slot):
            if loss_ev[slot] is None:
                return None
            loss_ev[slot].synchronize()
            return float(loss_host[slot])

        def e2e_step(i):
            batch, ev = pending['queue'].pop(0)
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            batch.data.record_stream(cur)
            for v in batch.domain.values():
                v.record_stream(cur)
            if pipelined:
                nxt, nev = pending['queue'][0]
                cur.wait_event(nev)
                nxt.data.record_stream(cur)
                for v in nxt.domain.values():
                    v.record_stream(cur)
                model.announce_next_batch(nxt)
            loss = model.training_step(batch, i)
            pending['queue'].append(prefetch(i + depth))   # issued once this step's graph is enqueued: off the launch critical path
            for p in params:
                p.grad = None
            loss.backward()
            flat = model.executor.last_flat_grad      # every p.grad is a view of this buffer
            trainer.reduce_gradients(flat)            # one NCCL all-reduce (N > 1); 1/world is folded into the Adam kernel
            trainer.step(flat)                        # fused Adam on the flat parameter buffer (one launch)
            if args.sync_loss:
                return float(loss.detach())           # blocking device -> host read of the step's result
            # device -> host read of EVERY step's loss through a pinned buffer; the host looks at it one step later (what a
            # logging callback does), so the GPU already has the next step queued while the host waits for this one
            slot = i & 1
            loss_host[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_ev[slot] = torch.cuda.Event()
            loss_ev[slot].record()
            return read_loss(slot ^ 1)

        for i in range(W):
            e2e_step(i)
        env.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(K):
            e2e_step(W + i)
        if not args.sync_loss:
            read_loss((W + K - 1) & 1)                # the last step's loss is read inside the timed region too
        f1.record()
        env.barrier()
        ms_e2e = f0.elapsed_time(f1)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- (3) per-kernel-family timing (eager, CUDA events around every C-ABI call) ------------
    ops.PROFILE = ops.KernelProfile()
    ex.ctx.overlap = False      # one stream: CUDA events around each call then time that call's kernels alone
    prof_steps = min(K, 5) if micro == 1 else 1
    for i in range(prof_steps):
        # a step is ~100 launches of 10-200 us kernels and Python needs ~40 us per call: without a head start the GPU
        # idles between calls and every event pair also times that idle gap.  A spin kernel in front of the step lets the
        # host queue the whole step before the first kernel runs, so the events bracket back-to-back kernels only.
        torch.cuda._sleep(40_000_000)
        model.fused_step(dev_batches[i % n_batches], args.laplacian)
        torch.cuda.synchronize()
    fam = ops.PROFILE.summary()
    ops.PROFILE = None
    ex.ctx.overlap = True

    ms_total, ms_e2e_m = env.max_over_ranks(ms_total, ms_e2e if ms_e2e is not None else 0.0)
    pts_per_step = world * b_per_gpu * ni
    out = {'value': pts_per_step * K / (ms_total / 1e3), 'ms_per_step': ms_total / K, 'steps': K, 'warmup': W,
           'loss': loss_value, 'clocks': sampler.summary(), 'graph': graph is not None, 'tail_in_graph': tail_in_graph,
           'pipelined': pipelined, 'h2d_bytes': h2d_bytes,
           # kernels of this library per optimizer step: every micro-batch's fused step + its input copy (+ the position
           # gather of the geometry branch), the gradient reset of an accumulated step, the two Adam launches
           'launches_per_step': (launches_per_micro + 1 + (1 if pipelined else 0)) * micro + 2 + (1 if micro > 1 else 0),
           'fallbacks': list(ops.FALLBACKS), 'b_micro': b_micro, 'micro': micro,
           'roofline': family_roofline(fam, prof_steps, env.peaks, engine_name) if rank == 0 else None,
           'dev_batches': dev_batches, 'labels': labels, 'spec': spec, 'model': model, 'trainer': trainer,
           'dp_tail': (None if world == 1 else 'nccl' if trainer.dp is None else ('multimem' if trainer.dp['multimem'] else 'p2p')),
           'dp_error': int(trainer.dp['epoch'][2]) if trainer.dp is not None else 0}
    if e2e:
        out['e2e_value'] = pts_per_step * K / (ms_e2e_m / 1e3)
        out['e2e_ms_per_step'] = ms_e2e_m / K
    return out


def extra_configs(env: Env) -> list:
    """The other BASELINE.json configs at this world size (see the module docstring).  Strong-scaled ones keep the total
    number of geometries of the config and split it over the ranks; large per-rank batches run as micro-batches."""
    world = env.world
    plan = []
    # config 3: PI-GANO duct_variable, batch 64 in total
    plan.append(dict(config=3, name='duct_pigano', shape=WORKLOADS['duct_pigano']['shape'], total=64, micro_geoms=64,
                     steps=20, scaling='strong'))
    # config 4: PI-GANO++ windbreaks, geometries per GPU swept (weak)
    for b in (1, 2, 4, 8):
        plan.append(dict(config=4, name='windbreaks_pigano_pp', shape=WORKLOADS['windbreaks_pigano_pp']['shape'], per_gpu=b,
                         micro_geoms=8, steps=10, scaling='weak'))
    # config 5: manufactured PIPN++ sweep, 256 geometries in total; NB = NI / 4, no observations
    # micro-batch sizes: as many geometries per launch as fit comfortably in HBM (~35 GB of jets at 1 M collocation points):
    # FPS runs one CTA per geometry, so few geometries per micro-batch leave it as the critical path
    for ni, mg, st in ((4096, 64, 10), (65536, 16, 3)) + (((262144, 4, 2),) if world >= 8 else ()):
        plan.append(dict(config=5, name='manufactured_pipn_pp', shape=dict(n_internal=ni, n_boundary=ni // 4, n_obs=0),
                         total=256, micro_geoms=mg, steps=st, scaling='strong'))
    out = []
    for p in plan:
        if 'total' in p:
            if p['total'] % world:
                continue
            b = p['total'] // world
        else:
            b = p['per_gpu']
        b_micro = min(b, p['micro_geoms'])
        while b % b_micro:
            b_micro -= 1
        micro = b // b_micro
        r = None
        try:
            r = run_workload(env, p['name'], p['shape'], b, p['steps'], 3, micro=micro, e2e=False, seed_base=7000)
            entry = {'config': p['config'], 'workload': f"{p['name']}: {WORKLOADS[p['name']]['what']}, {shape_text(p['shape'])} points, "
                                                        f"{b} geometries per GPU x {world} GPU(s)",
                     'global_batch': b * world, 'scaling': p['scaling'], 'value': r['value'], 'unit': 'points/s',
                     'ms_per_step': r['ms_per_step'], 'steps': r['steps'], 'warmup': r['warmup'],
                     'micro_batches_per_step': micro, 'cuda_graph': r['graph'], 'clocks': r['clocks'],
                     'fallbacks': r['fallbacks']}
            if r['roofline'] is not None:
                rf = r['roofline']
                entry['roofline'] = {k: rf[k] for k in ('bound', 'kernel', 'achieved', 'peak', 'unit', 'frac', 'share_of_step')}
                entry['roofline']['all_jet_gemms'] = rf['all_jet_gemms']
                entry['families_ms_per_step'] = rf['families_ms_per_step']
        except Exception as e:       # an auxiliary measurement must not cost the bench line
            entry = {'config': p['config'], 'workload': p['name'], 'error': f'{type(e).__name__}: {str(e)[:300]}'}
            torch.cuda.synchronize()
        out.append(entry)
        del r
        gc.collect()                 # model <-> executor is a reference cycle: free its graphs and buffers now
        torch.cuda.empty_cache()
    return out


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--engine', type=int, default=None,
                    help='jet GEMM engine: 2 = warp-specialised TMA + tcgen05 (default), 0 = the generic fp32 FFMA reference engine')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of replaying a CUDA graph')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the `configs` block (the other BASELINE configs)')
    ap.add_argument('--sync-loss', action='store_true', help='end-to-end leg: blocking float(loss) every step instead of the delayed read')
    ap.add_argument('--no-pipeline', action='store_true',
                    help='compute the set-abstraction geometry (FPS, ball query) of a batch inside its own step instead of one step ahead')
    ap.add_argument('--laplacian', default='reference', choices=['reference', 'true'])
    ap.add_argument('--config', default='abc_pipn_pp', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='geometries per GPU (default: the config\'s own)')
    ap.add_argument('--micro', type=int, default=1, help='micro-batches per optimizer step (gradient accumulation)')
    ap.add_argument('--n-internal', type=int, default=0,
                    help='collocation points per geometry (manufactured sweep: boundary = internal / 4)')
    args = ap.parse_args()
    config = args.config
    shape = dict(WORKLOADS[config]['shape'])
    b_per_gpu = args.batch or WORKLOADS[config]['batch']
    if args.n_internal:
        ratio = shape['n_boundary'] / shape['n_internal']
        obs_ratio = shape['n_obs'] / shape['n_internal']
        shape = dict(n_internal=args.n_internal, n_boundary=int(args.n_internal * ratio), n_obs=int(args.n_internal * obs_ratio))
    what = WORKLOADS[config]['what']
    if args.impl == 'reference':
        return run_reference(args, config, shape, b_per_gpu)

    from porous_cfd_b200 import ops
    env = Env(args)
    if args.engine is not None:
        ops.set_gemm_engine(args.engine)
    r = run_workload(env, config, shape, b_per_gpu, args.steps, args.warmup, micro=args.micro, e2e=True)
    headline_only = args.no_configs or config != 'abc_pipn_pp' or args.batch or args.n_internal or args.micro > 1

    # the auxiliary legs of the headline workload run BEFORE the other configs (N = 1 only, so every rank still enters
    # extra_configs together)
    ingest = None
    legs_last = os.environ.get('PCFD_BENCH_LEGS_LAST', '0') == '1'      # diagnosis of the capture issue in DESIGN.md section 8
    configs = None
    if legs_last and not headline_only:
        configs = extra_configs(env)
    if env.world == 1:
        try:
            ingest = ingest_leg(r['dev_batches'], r['labels'], shape['n_internal'], r['spec']['dims'], env.peaks)
            import tempfile
            with tempfile.TemporaryDirectory() as td:
                ingest['train_loop'] = train_loop_leg(config, r['dev_batches'], r['labels'], b_per_gpu, shape['n_internal'], td)
            if r['model'].executor.uses_geometry() and not args.no_graph:
                ingest['geometry_cache'] = geometry_cache_leg(r['model'], r['trainer'], r['dev_batches'], r['labels'], b_per_gpu,
                                                              shape['n_internal'], r['steps'])
        except Exception as e:      # an auxiliary measurement must not cost the bench line
            ingest = {'error': f'{type(e).__name__}: {e}'}
            torch.cuda.synchronize()

    if not legs_last:
        configs = None if headline_only else extra_configs(env)
    if env.rank != 0:
        if env.world > 1:
            env.dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline:
        pts, ms_cpu, threads = cpu_reference_step_rate(config, shape, steps=6, warmup=2, n_geom=2)
        cpu = {'value': pts, 'unit': 'points/s', 'cores': threads, 'kind': 'port', 'ms_per_step': ms_cpu,
               'sample': f'2 of the {b_per_gpu} geometries per step ({2 * shape["n_internal"]} collocation points), 6 timed steps '
                         f'after 2 warm-up (the whole batch on the CPU is what `--impl reference` times)'}

    line = {'metric': METRIC, 'value': r['value'], 'unit': 'points/s',
            'n_gpus': env.world, 'steps': r['steps'], 'warmup': r['warmup'], 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{config}: {what}, {shape_text(shape)} points, '
                                   f'{b_per_gpu} geometries per GPU', 'global_batch': env.world * b_per_gpu,
                       'laplacian': args.laplacian, 'dropout': 'on', 'optimizer': 'fused Adam in the step',
                       'parallelism': f'dp{env.world}', 'cuda_graph': r['graph'],
                       'collective': ('none (N = 1)' if env.world == 1 else
                                      {'nccl': 'one NCCL all-reduce of the flat gradient + Adam, ',
                                       'multimem': 'one kernel over NVSwitch multicast memory (pcfd_dp_adam_step: multimem.ld_reduce '
                                                   'reduce-scatter, Adam on the rank\'s slice, multimem.st all-gather of the parameters), ',
                                       'p2p': 'one kernel over NVLink peer memory (pcfd_dp_adam_step: peer loads, Adam on the rank\'s '
                                              'slice, peer stores), '}[r['dp_tail']] +
                                      ('captured in the step graph' if r['tail_in_graph'] else 'launched behind the step graph')),
                       'collective_barrier_timeouts': r['dp_error'],
                       'micro_batches_per_step': r['micro'],
                       'geometry': ('FPS / ball query of batch t+1 run as a parallel branch of step t (positions only; '
                                    'one batch worth per step)') if r['pipelined'] else 'inside the step',
                       'l2': 'per-step working set (jets + gradients ~1 GB) exceeds the 126 MB L2; 4 batches cycled, no flush'},
            'loss': r['loss'], 'clocks': r['clocks'],
            'e2e': {'value': r['e2e_value'], 'unit': 'points/s', 'ms_per_step': r['e2e_ms_per_step'], 'h2d_bytes_per_step': r['h2d_bytes'],
                    'd2h_bytes_per_step': 4, 'api': 'model.cuda_graph = True; model.training_step(model.transfer_batch_to_device(host_batch)); loss.backward(); ' +
                           ('FlatAdamTrainer.step(); float(loss)  [next batch prefetched on a copy stream]' if args.sync_loss else
                            'FlatAdamTrainer.step(); loss copied to pinned host memory every step and read by the host one step later  '
                            '[next batches prefetched on a copy stream]')},
            'gpu_launches': r['launches_per_step'] * r['steps'], 'tensor_core_fallbacks': r['fallbacks'],
            'roofline': r['roofline'], 'cpu_baseline': cpu, 'ingest': ingest, 'configs': configs}
    _emit(line)
    if env.world > 1:
        env.dist.destroy_process_group()


if __name__ == '__main__':
    main()
