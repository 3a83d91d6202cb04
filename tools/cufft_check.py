"""cuFFT strawman vs. the product path (BASELINE.json north_star: "cuFFT used only as a correctness and
performance oracle").  Runs on the GPU box:

    python tools/cufft_check.py [--planes 64] [--out gpurun_out/cufft_strawman.json]

1. builds tools/cufft_strawman.cu into tools/libcufft_strawman.so (own shared object, never part of
   libpsfr_b200.so);
2. correctness: its 40 x 40 psf_muse planes for two oracle PSDs x 35 wavelengths against the CPU oracle
   (relative 1e-9 on every pixel above 1e-6 of the peak) - and the product path on the same input;
3. performance: stage A + stage B (PSD -> structure function -> 35 PSFs -> 40 x 40 planes) of one chunk of
   `--planes` planes, cuFFT pipeline vs. psfr_structure_function + psfr_psf_cube, CUDA events, same box.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
SRC = os.path.join(ROOT, 'tools', 'cufft_strawman.cu')
LIB = os.path.join(ROOT, 'tools', 'libcufft_strawman.so')
LBDA = np.linspace(490, 930, 35)
N = 1280


def build():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(['nvcc', '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-shared',
                               '-Xcompiler', '-fPIC', SRC, '-lcufft', '-o', LIB])
    lib = ctypes.CDLL(LIB)
    P = ctypes.c_void_p
    lib.strawman_run.restype = ctypes.c_int
    lib.strawman_run.argtypes = [ctypes.c_int, ctypes.c_int, P, P, ctypes.c_int, P, P, ctypes.c_int,
                                 ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    return lib


def strawman(lib, psd, t_half, reps):
    psd = np.ascontiguousarray(psd)
    out = np.empty((psd.shape[0], LBDA.size, 40, 40))
    a, b = ctypes.c_double(), ctypes.c_double()
    rc = lib.strawman_run(N, psd.shape[0], psd.ctypes.data, t_half.ctypes.data, LBDA.size, LBDA.ctypes.data,
                          out.ctypes.data, reps, ctypes.byref(a), ctypes.byref(b))
    if rc:
        raise RuntimeError('strawman_run failed: %d' % rc)
    return out, a.value, b.value


def _oracle_block(args):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import psfr_oracle as orc
    psd, lam = args
    return orc.psf_muse(psd, lam)


def image_errors(got, ref):
    peak = ref.max(axis=(-1, -2), keepdims=True)
    sig = ref > 1e-6 * peak
    rel = np.where(sig, np.abs(got - ref) / np.where(sig, ref, 1.0), 0.0)
    return float(rel.max()), float((np.abs(got - ref) / peak).max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--planes', type=int, default=64)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'cufft_strawman.json'))
    a = ap.parse_args()
    import torch
    from joblib import Parallel, delayed
    import psfr_oracle as orc
    from muse_psfr_b200 import psfrec
    lib = build()
    psfrec.set_device(0)
    t_half = np.ascontiguousarray(np.fft.ifftshift(orc.telescope_otf(orc.pupil_mask(N / 4, N / 2, 0.14), N))[:, :N // 2 + 1])

    # ---- correctness: two oracle PSDs, 35 wavelengths, against the CPU oracle
    params = [(1.0, 0.7, 25.0), (0.55, 0.45, 12.0)]
    psd2 = np.stack([orc.simul_psd_wfm([g, 1 - g], (100, 10000), s, l0)[0] for s, g, l0 in params])
    cores = os.cpu_count() or 1
    blocks = [b for b in np.array_split(np.arange(LBDA.size), min(cores, LBDA.size)) if b.size]
    ref = np.empty((2, LBDA.size, 40, 40))
    for p in range(2):
        parts = Parallel(n_jobs=cores)(delayed(_oracle_block)((psd2[p], LBDA[b])) for b in blocks)
        ref[p] = np.concatenate(parts)
    got_cufft, _, _ = strawman(lib, psd2, t_half, 1)
    got_ours = np.stack([psfrec.psf_muse(psd2[p], LBDA) for p in range(2)])
    e_cufft, e_ours = image_errors(got_cufft, ref), image_errors(got_ours, ref)

    # ---- performance: one chunk of `planes` planes through stage A + stage B
    rng = np.random.default_rng(12345)
    n = a.planes
    seeing, GL, L0 = rng.uniform(0.4, 2.0, n), rng.uniform(0.3, 0.95, n), rng.uniform(9, 29, n)
    psd = np.concatenate([psfrec.simul_psd_wfm([GL[i], 1 - GL[i]], (100., 10000.), seeing[i], L0[i], verbose=False)
                          for i in range(n)])
    cube_cufft, ms_a, ms_b = strawman(lib, psd, t_half, a.reps)
    ctx = psfrec.get_context(max_planes=n, max_lambda=LBDA.size)
    d_psd = torch.from_numpy(psd).cuda()
    d_out = torch.empty((n, LBDA.size, 40, 40), dtype=torch.float64, device='cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    times = []
    for r in range(a.reps + 1):
        ctx.load_psd(d_psd, n)
        flush.zero_()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        ctx.structure_function(n)
        e1.record()
        ctx.psf_cube(n, 1, LBDA, d_out)
        e2.record()
        torch.cuda.synchronize()
        if r:
            times.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
    ours_a, ours_b = (float(np.mean([t[k] for t in times])) for k in (0, 1))
    agree = image_errors(d_out.cpu().numpy(), cube_cufft)
    psfs = n * LBDA.size
    res = {'planes': n, 'wavelengths': int(LBDA.size), 'dim': N, 'psfs_per_chunk': psfs,
           'stages': 'stage A (PSD -> structure function) + stage B (35 OTF -> PSF transforms, 40x40 resampling); PSD '
                     'synthesis, convolutions and the fit are outside both timings',
           'cufft': {'stage_a_ms': ms_a, 'stage_b_ms': ms_b, 'psf_per_s': psfs / ((ms_a + ms_b) * 1e-3),
                     'pipeline': 'shift kernel + cufftExecD2Z (batch = planes) + structure-function kernel; per plane: '
                                 'exp * OTF kernel (35 half planes) + cufftExecZ2D (batch 35) + sampling kernel',
                     'err_vs_oracle_rel': e_cufft[0], 'err_vs_oracle_peak': e_cufft[1]},
           'product': {'stage_a_ms': ours_a, 'stage_b_ms': ours_b, 'psf_per_s': psfs / ((ours_a + ours_b) * 1e-3),
                       'pipeline': 'psfr_structure_function + psfr_psf_cube (pruned row / column passes, default options)',
                       'err_vs_oracle_rel': e_ours[0], 'err_vs_oracle_peak': e_ours[1]},
           'speedup_stage_ab': (ms_a + ms_b) / (ours_a + ours_b),
           'product_vs_cufft_rel': agree[0], 'product_vs_cufft_peak': agree[1]}
    print(json.dumps(res, indent=1))
    assert e_cufft[0] < 1e-9 and e_ours[0] < 1e-9, 'strawman / product do not hold the 1e-9 bar against the oracle'
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, 'w') as f:
        json.dump(res, f, indent=1)
        f.write('\n')


if __name__ == '__main__':
    main()
