#!/usr/bin/env python
"""Benchmark of the BCNN VGG-16 448x448 train step (BASELINE.json metric) on N B200s, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--stage 2] [--batch 32] [--impl native|reference]

Prints ONE JSON line (rank 0).  `value` = device-timed images/s with inputs resident in HBM; `e2e` = the same step
through hawkeye_b200.train.Trainer.batch_training with pinned HOST inputs (H2D copy + loss/acc read-back inside the
timed region); `roofline` = the fused bilinear-pool forward (hk_bilinear_pool_fwd) against the measured HBM peak;
`cpu_baseline` = the oracle port of the same step on the host cores.  `--impl reference` times that CPU path only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch  # noqa: E402

K1_FWD_BYTES_PER_IMG = 1449984      # read X 512*196*4 + write Y 512*512*4
K1_BWD_BYTES_PER_IMG = 1851392
# dram__bytes_read.sum + dram__bytes_write.sum per launch of bcnn_gram_fwd_kernel from the ncu --set full captures
# summarised in profiles/gram_r1h_metrics.txt (tests/prof_bilinear.py 32 / 256)
K1_DRAM_TRAFFIC_B32 = 12.89e6
K1_DRAM_TRAFFIC_B256 = 102.85e6 + 208.88e6
VGG16_FWD_GFLOP_PER_IMG = 122.9
METRIC = '448x448 images/sec, BCNN VGG-16 train step (fwd+CE+bwd+grad all-reduce+SGD), device-timed, max over ranks'


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops_sustained', 1400.0), 'measured'
    return 6650.0, 1400.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def host_threads():
    """Threads for the CPU arm: all cores up to 32 (a batch-2 step stops scaling, and oversubscribes, beyond that)."""
    return max(1, min(os.cpu_count() or 1, 32))


def cpu_step_port(stage, B, threads, steps):
    """The oracle port of the reference step on the host cores (test/bench infrastructure, never the product)."""
    import detgen
    from oracle import hop_oracle as O
    torch.set_num_threads(threads)
    state = detgen.vgg_bcnn_state(O.VGG16_D, 200, seed=100)
    keys = None if stage == 2 else {'classifier.weight', 'classifier.bias'}
    x = torch.randn(B, 3, 448, 448, generator=torch.Generator().manual_seed(1234))
    labels = torch.randint(0, 200, (B,), generator=torch.Generator().manual_seed(1))
    bufs = {}

    def step():
        _, loss, grads = O.loss_and_grads(lambda xx, st: O.bcnn_forward(xx, st, stage), x, labels, state, keys)
        for k, g in grads.items():
            state[k], bufs[k] = O.sgd_momentum_step(state[k], g, bufs.get(k), 0.005, 0.9, 1e-5, k not in bufs)
        return loss

    step()  # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return B / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = host_threads()
    B = 2
    ips, dt = cpu_step_port(args.stage, B, threads, max(1, min(args.steps, 3)))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': 'img/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'BCNN VGG-16 stage {args.stage}, 448x448, 200 classes; CPU sample batch {B}'},
        'cpu_baseline': {'value': ips, 'unit': 'img/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{min(args.steps, 3)} timed steps of batch {B} (fwd+CE+bwd+SGD), torch-CPU oracle '
                                   f'port of the reference step, {threads} threads'},
        'e2e': {'value': ips, 'unit': 'img/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def time_bilinear_kernel(B, bwd=False, min_footprint=640 << 20, reps=4):
    """Average device time per call of hk_bilinear_pool_fwd (or _bwd): `reps` passes over a ring of buffer sets whose
    total footprint is >= 5x the 126 MB L2, launched back to back and bracketed by one CUDA-event pair on the launching
    stream.  Every call therefore reads inputs that are not L2-resident and runs while its predecessors' outputs are
    still being written back — the steady state of the kernel, with no separate flush kernel in the timed region."""
    from hawkeye_b200 import _lib
    per_set = B * (K1_BWD_BYTES_PER_IMG + 401408 if bwd else K1_FWD_BYTES_PER_IMG)
    nset = max(2, -(-min_footprint // per_set))
    xs = [torch.rand(B, 512, 14, 14, device='cuda') for _ in range(nset)]
    ys = [torch.empty(B, 512 * 512, device='cuda') for _ in range(nset)]
    if bwd:
        for y in ys:
            y.normal_()
        dxs = [torch.empty_like(x) for x in xs]
    name = 'hk_bilinear_pool_bwd' if bwd else 'hk_bilinear_pool_fwd'
    nb = _lib.query(name + '_workspace_bytes', B, 512, 196)
    ws = torch.empty(nb, dtype=torch.uint8, device='cuda')
    s = _lib.stream_ptr()

    def call(i):
        if bwd:
            _lib.call(name, xs[i], ys[i], dxs[i], B, 512, 196, ws, nb, s)
        else:
            _lib.call(name, xs[i], ys[i], None, B, 512, 196, ws, nb, s)

    for i in range(min(nset, 3)):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(4e6))        # ~2 ms of device-side spin so the host enqueues ahead of the GPU (no launch gaps)
    e0.record()
    for _ in range(reps):
        for i in range(nset):
            call(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (reps * nset)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--stage', type=int, default=2)
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--impl', default='native')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='profiling runs only')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)

    import torch.distributed as dist
    from hawkeye_b200 import _lib, engine
    from hawkeye_b200.config import load_config
    from hawkeye_b200.train import Trainer

    rank, local, world = engine.init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    cfg = load_config(os.path.join(ROOT, 'configs', f'BCNN_S{args.stage}.yaml'))
    torch.manual_seed(0)                                   # random-init weights, reference initialisers
    tr = Trainer(cfg, dataloaders={})
    tr.model.train()
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, 3, 448, 448, generator=g).pin_memory()
    y_host = torch.randint(0, 200, (B,), generator=g).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step_resident():
        out = tr.model(x_dev)
        loss = tr.criterion(out, y_dev)
        tr.optimizer.zero_grad()
        loss.backward()
        tr.allreduce.finish()
        tr.optimizer.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        loss = step_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    value = B * world * args.steps / (ms * 1e-3)

    # end to end through the public Trainer API with host inputs
    data = {'img': x_host, 'label': y_host}
    if args.no_e2e:
        ms_e2e = float('nan')
    else:
        for _ in range(2):
            tr.batch_training(data)
        ms_e2e = timed(lambda: tr.batch_training(data), args.steps)
    e2e = B * world * args.steps / (ms_e2e * 1e-3)
    final_loss = float(loss.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, tf_peak, which = measured_peaks()
    if args.no_e2e:
        t_avg, t256, t1024 = float('nan'), float('nan'), float('nan')
    else:
        t_avg = time_bilinear_kernel(32)
        t256 = time_bilinear_kernel(256)
        t1024 = time_bilinear_kernel(1024)
    ach = 32 * K1_FWD_BYTES_PER_IMG / t_avg / 1e9
    ach256 = 256 * K1_FWD_BYTES_PER_IMG / t256 / 1e9
    ach1024 = 1024 * K1_FWD_BYTES_PER_IMG / t1024 / 1e9
    flops_img = VGG16_FWD_GFLOP_PER_IMG * (3.0 if args.stage == 2 else 1.0) * 1e9
    conv_tf = flops_img * B * args.steps / (ms * 1e-3) / 1e12
    line = {
        'metric': METRIC, 'value': value, 'unit': 'img/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'tf32 (fp32 storage, fp32 accumulate)', 'data': 'synthetic',
        'config': {'workload': f'BCNN VGG-16 stage {args.stage}, 448x448, batch {B}/GPU, 200 classes, SGD momentum',
                   'global_batch': B * world, 'parallelism': f'dp{world}',
                   'l2': 'per-step working set (7.7 GB activations) >> 126 MB L2; pool microbench rotates through buffer '
                         'sets totalling >= 640 MB (5x L2), so every launch reads cold inputs',
                   'final_loss': final_loss},
        'clocks': clocks,
        'e2e': {'value': e2e, 'unit': 'img/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': x_host.numel() * 4 + y_host.numel() * 8, 'd2h_bytes_per_step': 8},
        'gpu_launches': launches,
        'roofline': {'kernel': 'hk_bilinear_pool_fwd = bcnn_gram_fwd_kernel<512> (one launch: Gram + sqrt + L2 normalise), '
                               'B=32 (the per-GPU batch of this workload), C=512, HW=196',
                     'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
                     'traffic': K1_DRAM_TRAFFIC_B32, 'peak_source': which, 'us_per_launch': t_avg * 1e6,
                     'algorithmic_bytes_per_launch': 32 * K1_FWD_BYTES_PER_IMG,
                     'note': 'B=32 moves 46 MB in ~20 us: launch/fill latency bound, and the 33.5 MB output stays in the '
                             '126 MB L2 (ncu: 12.9 MB of DRAM traffic per launch); the streaming regime is b256/b1024',
                     'b256': {'achieved': ach256, 'frac': ach256 / hbm_peak, 'us_per_launch': t256 * 1e6,
                              'traffic': K1_DRAM_TRAFFIC_B256},
                     'b1024': {'achieved': ach1024, 'frac': ach1024 / hbm_peak, 'us_per_launch': t1024 * 1e6}},
        'roofline_conv': {'bound': 'tensor', 'achieved': conv_tf, 'unit': 'TFLOP/s (tf32, whole step incl. non-conv time)',
                          'peak': tf_peak / 2, 'frac': conv_tf / (tf_peak / 2),
                          'note': 'peak = measured bf16 sustained / 2 (tf32 runs at half the bf16 rate)'},
    }
    if not args.no_cpu_baseline:
        threads = host_threads()
        ips, dt = cpu_step_port(args.stage, 2, threads, 2)
        line['cpu_baseline'] = {'value': ips, 'unit': 'img/s', 'cores': threads, 'kind': 'port',
                                'sample': f'2 timed steps of batch 2 ({dt:.2f} s/step), torch-CPU oracle port of the '
                                          f'reference BCNN step (fwd+CE+bwd+SGD), {threads} threads'}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
