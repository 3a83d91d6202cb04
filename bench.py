"""Benchmark of the PSF-reconstruction hot path (BASELINE.json metric: PSFs/sec at dim 1280,
PSD -> PSF -> Moffat fit).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--draws D] [--impl reference]

Workload (config.workload): BASELINE.json configs[3] - the sweep of random (seeing, GL, L0,
Cn2 profile) draws x 35 wavelengths, `--draws` (default 4096) draws per GPU per step, weak
scaling (rank r uses the same distributions with seed 12345 + r).  One step = the whole path
for every draw: PSD synthesis, structure function, 35 pruned OTF -> PSF transforms, resample,
the two Moffat convolutions and the Moffat fit.

One JSON line on rank 0; see DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
LBDA = np.linspace(490, 930, int(os.environ.get('PSFR_BENCH_NLAM', '35')))   # 35 = BASELINE config; env override for tuning only
N = 1280
# canonical algorithmic bytes per PSF (two-pass real-input 2-D FFT, FP64, no
# pruning / symmetry / L2 credit): stage B 32 N^2 + stage A 40 N^2 / nlam
BYTES_STAGE_B = 32 * N * N
BYTES_PER_PSF = BYTES_STAGE_B + 40 * N * N / LBDA.size


def draws_for(rank, nd):
    rng = np.random.default_rng(12345 + rank)
    seeing = rng.uniform(0.4, 2.0, nd)
    GL = rng.uniform(0.3, 0.95, nd)
    L0 = rng.uniform(9, 29, nd)
    h = np.stack([rng.uniform(50, 500, nd), rng.uniform(5000, 15000, nd)], axis=1)
    return seeing, GL, L0, h


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower() == 'active' for r in self.rows if len(r) > 3 + k)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None,
                'sm_max_mhz': float(self.rows[0][1]) if self.rows[0][1].replace('.', '').isdigit() else None,
                'power_w_max': max(float(r[2]) for r in self.rows if r[2].replace('.', '').isdigit()) if self.rows else None,
                'samples': len(self.rows), 'reasons': reasons}


# ------------------------------------------------------------------------------- CPU arm
def _cpu_draw(args):
    """One draw through the CPU oracle (the reference's algorithm), single-threaded numpy."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import psfr_oracle as orc
    seeing, GL, L0, h, lbda = args
    orc.compute_psf(lbda, seeing, GL, L0, h=tuple(h))
    return len(lbda)


def cpu_throughput(ndraw, nlam, cores):
    """PSFs/s of the oracle driven like the reference (joblib over draws, psfrec.py:1082)."""
    from joblib import Parallel, delayed
    seeing, GL, L0, h = draws_for(0, ndraw)
    lbda = LBDA[:nlam]
    jobs = [(seeing[i], GL[i], L0[i], h[i], lbda) for i in range(ndraw)]
    with Parallel(n_jobs=cores) as par:
        par(delayed(time.sleep)(0.01) for _ in range(cores))     # start the workers outside the timing
        t0 = time.time()
        done = par(delayed(_cpu_draw)(j) for j in jobs)
        dt = time.time() - t0
    return sum(done) / dt, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: one full draw (35 wavelengths, like the GPU arm) per core, ~12 s per step
    nlam, ndraw = LBDA.size, cores
    times, vals = [], []
    for step in range(args.warmup + args.steps):
        v, dt = cpu_throughput(ndraw, nlam, cores)
        if step >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = float(np.mean(vals))
    sample = ('%d draws of the config-4 sweep (seed 12345) x %d wavelengths per step, one joblib worker per host core; '
              'numpy oracle port of psfrec.py (the reference itself needs astropy/mpdaf, absent here)' % (ndraw, nlam))
    line = {'impl': 'reference', 'metric': 'PSFs/sec (dim 1280, PSD->PSF->Moffat fit)', 'value': value,
            'unit': 'PSF/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': float(np.mean(times) * 1e3), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'configs[3] bounded sample: ' + sample},
            'cpu_baseline': {'value': value, 'unit': 'PSF/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'PSF/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from muse_psfr_b200 import _lib, psfrec

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at VERSION level
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    psfrec.set_device(local)
    dev = torch.device('cuda', local)
    nd, nlam = args.draws, LBDA.size
    seeing, GL, L0, h = draws_for(rank, nd)
    ctx = psfrec.get_context(max_planes=args.max_planes, max_lambda=nlam, device=local)
    stream = torch.cuda.current_stream().cuda_stream
    # tuning knobs (None = the library defaults, which is what the reported line uses)
    if args.grade is not None:
        ctx.set_option(_lib.OPT_EXP_GRADE, args.grade)
    if args.f32_rows is not None:
        ctx.set_option(_lib.OPT_F32_ROWS, args.f32_rows)
    if args.exp_cut is not None:
        ctx.set_option(_lib.OPT_EXP_CUT, args.exp_cut)
    if args.row_kernel is not None:
        ctx.set_option(_lib.OPT_ROW_KERNEL, args.row_kernel)

    # ---- device-resident arm: inputs (draw records, tables) and outputs live in HBM
    recs = psfrec.draw_records(seeing, GL, L0, h)
    d_recs = torch.from_numpy(recs).to(dev)
    dirs = psfrec.direction_perf(1)
    pos = psfrec._lgs_positions(False)
    d_cube = torch.empty((nd, nlam, 40, 40), dtype=torch.float64, device=dev)
    d_fit = torch.empty((nd, nlam, _lib.FIT_NPAR), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        flush.zero_()                       # L2 flush between timed iterations
        ctx.compute_batch(d_recs, dirs, pos, LBDA, out_cube=d_cube, out_fit=d_fit, stream=stream)

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches()
    hot_ms, hot_n, hot_psfs = 0.0, 0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
        ms, n, psfs = ctx.last_hot_timing()     # CUDA events around every launch of the row kernel
        hot_ms, hot_n, hot_psfs = hot_ms + ms, hot_n + n, hot_psfs + psfs
    e1.record()
    barrier()
    launches = ctx.kernel_launches() - launches0
    t_dev = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_total = float(t_dev.item())

    # ---- end-to-end arm: the public Python API with HOST buffers (pinned outputs)
    h_cube = torch.empty((nd, nlam, 40, 40), dtype=torch.float64, pin_memory=True)
    h_fit = torch.empty((nd, nlam, _lib.FIT_NPAR), dtype=torch.float64, pin_memory=True)

    def step_e2e():
        psfrec.compute_psf_batch(LBDA, seeing, GL, L0, h=h, out_cube=h_cube, out_fit=h_fit,
                                 device=local, max_planes=args.max_planes, stream=stream)

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)

    # sanity on the result of the last step (loss-like read-back): fitted FWHM must be finite
    fw = h_fit[:, :, _lib.FIT_FWHM].numpy() * 0.2
    finite = bool(np.isfinite(fw).all())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    psfs_step = nd * nlam
    value = world * psfs_step * args.steps / (ms_total * 1e-3)
    e2e = world * psfs_step * args.steps / float(t_e2e.item())
    peak, peak_src = measured_peak()
    hot_avg_ms = hot_ms / max(hot_n, 1)
    achieved = (BYTES_STAGE_B * hot_psfs / max(hot_n, 1)) / (hot_avg_ms * 1e-3) / 1e9 if hot_n else None
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'hot_rows_traffic.json')) as f:
            traffic = json.load(f).get('dram_bytes_per_launch')
    except Exception:
        pass
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, dt = cpu_throughput(cores, nlam, cores)
        cpu = {'value': v, 'unit': 'PSF/s', 'cores': cores, 'kind': 'port',
               'sample': '%d draws of the same sweep x %d wavelengths, joblib over draws (one worker per core), '
                         'numpy oracle port of psfrec.py (%.1f s)' % (cores, nlam, dt)}
    line = {
        'metric': 'PSFs/sec (dim 1280, PSD->PSF->Moffat fit)', 'value': value, 'unit': 'PSF/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'configs[3]: sweep of random (seeing, GL, L0, Cn2 profile) draws x 35 wavelengths '
                               '490-930 nm, dim 1280, npsflin 1, 4 LGS; %d draws per GPU per step' % nd,
                   'draws_per_gpu': nd, 'wavelengths': nlam, 'chunk_planes': args.max_planes,
                   'l2': '256 MB buffer rewritten before every step (inside the timed region); per-chunk '
                         'working set ~4 GB >> 126 MB L2',
                   'exp_cut': 'OTF entries below exp(-64) = 1.6e-28 of the peak are flushed to zero (DESIGN.md 3.7)',
                   'graded_precision': 'row pairs entirely below exp(-%g) of the OTF peak are evaluated and transformed '
                                       'in FP32, blocks entirely below exp(-%g) use the FP32 exp; everything else '
                                       'FP64 (psfr.h PSFR_OPT_F32_ROWS / PSFR_OPT_EXP_GRADE; parity tests hold the '
                                       '1e-9 PSF bar)' % (25.0 if args.f32_rows is None else args.f32_rows,
                                                          20.0 if args.grade is None else args.grade),
                   'results_finite': finite},
        'e2e': {'value': e2e, 'unit': 'PSF/s',
                'h2d_bytes_per_step': int(recs.nbytes + dirs.nbytes + pos.nbytes + LBDA.nbytes),
                'd2h_bytes_per_step': int(h_cube.numel() * 8 + h_fit.numel() * 8)},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'kernel': 'stage-B row kernel (group_rows_kernel, or hot_rows_kernel with --row-kernel 1: exp(-c D)*OTF + 1280-pt FFT, pruned)',
                     'achieved': achieved, 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                     'frac': (achieved / peak) if achieved else None, 'traffic': traffic,
                     'algorithmic_bytes_per_psf': BYTES_STAGE_B, 'psfs_per_launch': hot_psfs / max(hot_n, 1),
                     'avg_launch_ms': hot_avg_ms,
                     'pipeline_frac': (value / world) * BYTES_PER_PSF / 1e9 / peak,
                     'note': 'canonical FULL-GRID bytes (SURVEY 8d); the kernel is pruned + lambda-batched and moves far '
                             'fewer bytes, so frac can exceed 1 - see DESIGN.md'},
        'clocks': sampler.summary(),
    }
    if cpu:
        line['cpu_baseline'] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--draws', type=int, default=4096, help='draws per GPU per step')
    ap.add_argument('--max-planes', type=int, default=64, dest='max_planes')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--grade', type=float, default=None, help='PSFR_OPT_EXP_GRADE override (tuning)')
    ap.add_argument('--f32-rows', type=float, default=None, dest='f32_rows', help='PSFR_OPT_F32_ROWS override (tuning)')
    ap.add_argument('--row-kernel', type=int, default=None, dest='row_kernel', help='PSFR_OPT_ROW_KERNEL override (tuning)')
    ap.add_argument('--exp-cut', type=float, default=None, dest='exp_cut', help='PSFR_OPT_EXP_CUT override (tuning)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
