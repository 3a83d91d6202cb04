"""Hardware parity sweep of the pruned stage-B path (underflow cut + graded precision; the same with
both grades off; and with the cut off as well) against the CPU oracle, over the BASELINE configs at
their stated sizes.

    python tools/parity_sweep.py [--out gpurun_out/parity_sweep.json] [--draws 64] [--configs 4,2,3,5]

Runs on the GPU box; the oracle (oracle/psfr_oracle.py, the checker) runs on all host cores through
joblib.  Per config and option set it reports
    img_rel   largest pointwise relative error over the pixels above 1e-6 of their plane's peak
    img_peak  largest |difference| / peak
    fwhm_rel, beta_rel   largest relative error of the fitted FWHM / beta
    not_converged        number of fits with PSFR_FIT_ITER < 0
against the bars of BASELINE.json (1e-9 images, 1e-5 fit).  tests/test_gpu_parity.py runs a 16-draw
slice of the config-4 part.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

LBDA35 = np.linspace(490, 930, 35)
FLOOR = 1e-6
OPTION_SETS = {'default': {}, 'allfp64': {'grade': 1e30, 'f32_rows': 1e30},
               'allfp64_nocut': {'grade': 1e30, 'f32_rows': 1e30, 'exp_cut': 1000.0}}


def config4_draws(n, corners=True):
    """First n draws of bench.py's config-4 stream (seed 12345) plus the 8 corners of the sweep's
    (seeing, L0, GL) box at h = (100, 10000)."""
    rng = np.random.default_rng(12345)
    full = 4096
    seeing, GL, L0 = rng.uniform(0.4, 2.0, full), rng.uniform(0.3, 0.95, full), rng.uniform(9, 29, full)
    h = np.stack([rng.uniform(50, 500, full), rng.uniform(5000, 15000, full)], axis=1)
    seeing, GL, L0, h = seeing[:n], GL[:n], L0[:n], h[:n]
    if corners:
        cs = np.array([(s, g, l) for s in (0.4, 2.0) for l in (9.0, 29.0) for g in (0.3, 0.95)])
        seeing = np.concatenate([seeing, cs[:, 0]])
        GL = np.concatenate([GL, cs[:, 1]])
        L0 = np.concatenate([L0, cs[:, 2]])
        h = np.concatenate([h, np.tile([100.0, 10000.0], (8, 1))])
    return seeing, GL, L0, h


def _oracle_job(job):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import psfr_oracle as orc
    lbda, seeing, GL, L0, h, kw = job
    res, cube = orc.compute_psf(lbda, seeing, GL, L0, h=tuple(h), **kw)
    return cube, res['fwhm'] / 0.2, res['n']


def oracle_many(jobs, cores=None):
    from joblib import Parallel, delayed
    cores = cores or os.cpu_count() or 1
    return Parallel(n_jobs=cores)(delayed(_oracle_job)(j) for j in jobs)


def compare(cube, fit, ref_cube, ref_fwhm, ref_n, fit_fwhm=5, fit_n=4, fit_iter=7):
    """Error figures of GPU results [n, nl, 40, 40] / [n, nl, 16] against the oracle's."""
    cube, ref_cube = np.asarray(cube), np.asarray(ref_cube)
    peak = ref_cube.max(axis=(-1, -2), keepdims=True)
    diff = np.abs(cube - ref_cube)
    sig = ref_cube > FLOOR * peak
    rel = np.where(sig, diff / np.where(sig, ref_cube, 1.0), 0.0)
    return {'img_rel': float(rel.max()), 'img_peak': float((diff / peak).max()),
            'fwhm_rel': float(np.abs(fit[..., fit_fwhm] / ref_fwhm - 1).max()),
            'beta_rel': float(np.abs(fit[..., fit_n] / ref_n - 1).max()),
            'not_converged': int((fit[..., fit_iter] < 0).sum()), 'finite': bool(np.isfinite(cube).all()),
            'planes': int(np.prod(cube.shape[:-2]))}


def set_options(ctx, opts):
    from muse_psfr_b200 import _lib
    info = ctx.info()
    ctx.set_option(_lib.OPT_EXP_GRADE, opts.get('grade', info['exp_grade']))
    ctx.set_option(_lib.OPT_F32_ROWS, opts.get('f32_rows', info['f32_rows']))
    ctx.set_option(_lib.OPT_EXP_CUT, opts.get('exp_cut', info['exp_cut']))
    return info


def restore_options(ctx, info):
    from muse_psfr_b200 import _lib
    ctx.set_option(_lib.OPT_EXP_GRADE, info['exp_grade'])
    ctx.set_option(_lib.OPT_F32_ROWS, info['f32_rows'])
    ctx.set_option(_lib.OPT_EXP_CUT, info['exp_cut'])


def gpu_batch(psfrec, opts, lbda, seeing, GL, L0, dim=1280, **kw):
    ctx = psfrec.get_context(max_lambda=len(lbda), dim=dim)
    info = set_options(ctx, opts) if dim == 1280 else None
    try:
        fit, cube = psfrec.compute_psf_batch(lbda, seeing, GL, L0, dim=dim, **kw)
    finally:
        if info:
            restore_options(ctx, info)
    return fit, cube


def sweep_config4(psfrec, ndraw=64, corners=True, option_sets=OPTION_SETS, cores=None):
    seeing, GL, L0, h = config4_draws(ndraw, corners)
    ref = oracle_many([(LBDA35, seeing[i], GL[i], L0[i], h[i], {}) for i in range(seeing.size)], cores)
    ref_cube = np.stack([r[0] for r in ref])
    ref_fw, ref_n = np.stack([r[1] for r in ref]), np.stack([r[2] for r in ref])
    out = {'draws': int(seeing.size), 'wavelengths': 35,
           'what': 'first %d draws of the seed-12345 stream%s x 35 wavelengths' % (
               ndraw, ' + the 8 corners of (seeing 0.4/2.0, L0 9/29, GL 0.3/0.95)' if corners else '')}
    for name, opts in option_sets.items():
        fit, cube = gpu_batch(psfrec, opts, LBDA35, seeing, GL, L0, h=h)
        out[name] = compare(cube, fit, ref_cube, ref_fw, ref_n)
        if corners:
            out[name]['corners'] = compare(cube[-8:], fit[-8:], ref_cube[-8:], ref_fw[-8:], ref_n[-8:])
    return out


def sparta_rows(n=30):
    sys.path.insert(0, ROOT)
    import bench
    return bench.sparta_rows(n)


def sweep_config2(psfrec, option_sets=OPTION_SETS, cores=None):
    """30 rows (3 in three-LGS mode), mean of the lasers, time mean + refit."""
    import psfr_oracle as orc
    jobs = psfrec.select_sparta_rows(sparta_rows(30))
    s, g, l0, three = (np.array(c) for c in list(zip(*jobs))[:4])
    ref = oracle_many([(LBDA35, s[i], g[i], l0[i], (100, 10000), {'three_lgs_mode': bool(three[i])})
                       for i in range(len(jobs))], cores)
    ref_cube = np.stack([r[0] for r in ref])
    ref_fw, ref_n = np.stack([r[1] for r in ref]), np.stack([r[2] for r in ref])
    ref_mean = ref_cube.mean(axis=0)
    mfit = orc.fit_psf_cube(LBDA35, ref_mean)
    out = {'rows': len(jobs), 'three_lgs_rows': int(three.sum()), 'wavelengths': 35}
    for name, opts in option_sets.items():
        fit = np.empty((len(jobs), 35, 16))
        cube = np.empty((len(jobs), 35, 40, 40))
        for mode in (False, True):
            sel = np.where(three == mode)[0]
            if sel.size:
                f, c = gpu_batch(psfrec, opts, LBDA35, s[sel], g[sel], l0[sel], three_lgs_mode=bool(mode))
                fit[sel], cube[sel] = f, c
        out[name] = compare(cube, fit, ref_cube, ref_fw, ref_n)
        mean, fmean = np.empty((35, 40, 40)), np.empty((35, 16))
        psfrec.get_context().mean_refit(len(jobs), 35, cube, mean, fmean)
        out[name]['mean'] = compare(mean[None], fmean[None], ref_mean[None], mfit['fwhm'][None] / 0.2, mfit['n'][None])
    return out


def _split(lbda, parts):
    return [c for c in np.array_split(lbda, min(parts, len(lbda))) if c.size]


def sweep_config3(psfrec, option_sets=OPTION_SETS, cores=None):
    """npsflin = 3, three-LGS mode, all 35 wavelengths (the oracle runs wavelength blocks in parallel)."""
    cores = cores or os.cpu_count() or 1
    blocks = _split(LBDA35, cores)
    ref = oracle_many([(b, 1.0, 0.7, 25.0, (100, 10000), {'npsflin': 3, 'three_lgs_mode': True}) for b in blocks], cores)
    ref_cube = np.concatenate([r[0] for r in ref])[None]
    ref_fw, ref_n = np.concatenate([r[1] for r in ref])[None], np.concatenate([r[2] for r in ref])[None]
    out = {'directions': 9, 'wavelengths': 35}
    for name, opts in option_sets.items():
        fit, cube = gpu_batch(psfrec, opts, LBDA35, [1.0], [0.7], [25.0], npsflin=3, three_lgs_mode=True)
        out[name] = compare(cube, fit, ref_cube, ref_fw, ref_n)
    return out


def sweep_config5(psfrec, nlam=10, cores=None):
    """dim 2560: `nlam` of the 100 wavelengths (evenly spread, both ends included)."""
    cores = cores or os.cpu_count() or 1
    lam100 = np.linspace(490, 930, 100)
    idx = np.unique(np.round(np.linspace(0, 99, nlam)).astype(int))
    blocks = _split(lam100[idx], cores)
    ref = oracle_many([(b, 1.0, 0.7, 25.0, (100, 10000), {'dim': 2560}) for b in blocks], cores)
    ref_cube = np.concatenate([r[0] for r in ref])
    ref_fw, ref_n = np.concatenate([r[1] for r in ref]), np.concatenate([r[2] for r in ref])
    fit, cube = gpu_batch(psfrec, {}, lam100, [1.0], [0.7], [25.0], dim=2560)     # all 100 on the GPU
    out = {'wavelengths_checked': int(idx.size), 'wavelengths_run': 100,
           'default': compare(cube[:, idx], fit[:, idx], ref_cube[None], ref_fw[None], ref_n[None])}
    out['default']['not_converged_of_100'] = int((fit[..., 7] < 0).sum())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'parity_sweep.json'))
    ap.add_argument('--draws', type=int, default=64)
    ap.add_argument('--configs', default='4,2,3,5')
    a = ap.parse_args()
    from muse_psfr_b200 import psfrec
    psfrec.set_device(0)
    res = {'cores': os.cpu_count(), 'bars': {'img_rel': 1e-9, 'fit_rel': 1e-5}, 'floor': FLOOR}
    for c in a.configs.split(','):
        t0 = time.time()
        if c == '4':
            res['config4'] = sweep_config4(psfrec, a.draws)
        elif c == '2':
            res['config2'] = sweep_config2(psfrec)
        elif c == '3':
            res['config3'] = sweep_config3(psfrec)
        elif c == '5':
            res['config5'] = sweep_config5(psfrec)
        res['config' + c]['seconds'] = round(time.time() - t0, 1)
        print('config', c, json.dumps(res['config' + c]), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, 'w') as f:
        json.dump(res, f, indent=1)
        f.write('\n')


if __name__ == '__main__':
    main()
