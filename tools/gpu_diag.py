"""Stage-by-stage comparison of the CUDA path with the oracle on one GPU.

Development aid: unlike the pytest suite it never stops at the first mismatch, so one
GPU round trip reports the state of every stage.  Usage (on the GPU box):
    python tools/gpu_diag.py > gpurun_out/diag.log 2>&1
"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

import psfr_oracle as orc  # noqa: E402
from muse_psfr_b200 import _lib, psfrec  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
LBDA35 = np.linspace(490, 930, 35)


def stat(name, got, ref, floor=1e-300):
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    d = np.abs(got - ref)
    scale = np.abs(ref).max()
    rel = d / np.maximum(np.abs(ref), floor)
    sig = np.abs(ref) > 1e-6 * scale
    print('%-34s max|d|/max %.3e   max rel (|ref|>1e-6 max) %.3e   nan %d' % (
        name, d.max() / scale, rel[sig].max() if sig.any() else 0.0, int(np.isnan(got).sum())), flush=True)


def section(fn):
    print('\n=== %s' % fn.__name__, flush=True)
    t = time.time()
    try:
        fn()
    except Exception:
        traceback.print_exc()
    print('--- %.2f s' % (time.time() - t), flush=True)


def s_otf():
    ctx = psfrec.get_context()
    t = ctx.get_otf()
    pup = orc.pupil_mask(320, 640, 0.14)
    ref = orc.telescope_otf(pup, 1280)
    stat('telescope OTF half-plane', t[:641], ref[:641])
    stat('telescope OTF (transposed ref)', t[:641], ref.T[:641])
    print('pad row max', np.abs(t[641]).max(), 'centre*N^2', t[640, 640] * 1280 ** 2, 'zeros', int((t[:641] == 0).sum()))


PSD1 = {}


def s_psd():
    for tag, args, kw in [('cfg1', ([0.7, 0.3], (100, 10000), 1.0, 25.), dict(npsflin=1)),
                          ('cfg3', ([0.7, 0.3], (100, 10000), 1.0, 25.), dict(npsflin=3, three_lgs_mode=True)),
                          ('float-h', ([0.55, 0.45], (150.5, 12000.), 0.63, 12.5), dict(npsflin=2))]:
        got = psfrec.simul_psd_wfm(*args, verbose=False, **kw)
        ref = orc.simul_psd_wfm(*args, **kw)
        stat('psd ' + tag, got, ref)
        stat('psd AO zone ' + tag, got[:, 600:680, 600:680], ref[:, 600:680, 600:680])
        print('   zero pattern equal:', np.array_equal(got == 0, ref == 0), ' sum rel err %.2e' % abs(got.sum() / ref.sum() - 1))
        if tag == 'cfg1':
            PSD1['psd'] = ref


def s_structure():
    psd = PSD1.get('psd')
    if psd is None:
        psd = orc.simul_psd_wfm([0.7, 0.3], (100, 10000), 1.0, 25.)
        PSD1['psd'] = psd
    ctx = psfrec.get_context()
    ctx.load_psd(np.ascontiguousarray(psd[0]), 1)
    ctx.structure_function(1)
    d = ctx.get_structure_function(0)
    ref = orc.structure_function_unit(psd[0])
    stat('D_unit (transposed half-plane)', d[:641], ref.T[:641])
    stat('D_unit (if not transposed)', d[:641], ref[:641])
    print('centre', d[640, 640], 'pad', np.abs(d[641]).max(), 'ref max', ref.max())


def s_full_psf():
    psd = PSD1['psd'][0]
    pup = orc.pupil_mask(320, 640, 0.14)
    for lb in (490e-9, 930e-9):
        got = psfrec.psd_to_psf(psd, pup, 8, lb, samp=2)
        ref = orc.psd_to_psf(psd, pup, 8, lb)
        stat('psd_to_psf %.0f nm' % (lb * 1e9), got, ref)
        print('   sum', got.sum(), 'argmax', np.unravel_index(got.argmax(), got.shape), 'min', got.min(), ref.min())


def s_psf_muse():
    g = np.load(os.path.join(GOLD, 'ref_config1.npz'))
    got = psfrec.psf_muse(PSD1['psd'][0], LBDA35)
    stat('psf_muse cfg1 (35 lambda)', got, g['psf_muse'])
    for i in (0, 17, 34):
        stat('   plane %d' % i, got[i], g['psf_muse'][i])
    g3 = np.load(os.path.join(GOLD, 'ref_config3.npz'))
    psd3 = orc.simul_psd_wfm([0.7, 0.3], (100, 10000), 1.0, 25., npsflin=3, three_lgs_mode=True)
    got3 = psfrec.psf_muse(psd3, g3['lbda'])
    stat('psf_muse cfg3 (9 dirs)', got3, g3['psf_muse'])


def s_convolve_fit():
    g, go = np.load(os.path.join(GOLD, 'ref_config1.npz')), np.load(os.path.join(GOLD, 'oracle_config1.npz'))
    conv = psfrec.convolve_final_psf(LBDA35, 1.0, 0.7, 25., g['psf_muse'])
    stat('convolve_final_psf', conv, go['conv'])
    tab = psfrec.fit_psf_cube(LBDA35, go['conv'])
    stat('fit fwhm (oracle conv in)', tab['fwhm'][:, 0], go['fwhm'])
    stat('fit n', tab['n'], go['n'])
    stat('fit centre', tab['center'], go['center'])
    ctx = psfrec.get_context()
    fit = np.empty((35, _lib.FIT_NPAR))
    ctx.moffat_fit(35, 40, 40, np.ascontiguousarray(go['conv']), fit)
    print('   LM iterations', fit[:, _lib.FIT_ITER].astype(int).tolist())
    print('   chisq', fit[[0, 17, 34], _lib.FIT_CHISQ])


def s_compute_psf():
    go = np.load(os.path.join(GOLD, 'oracle_config1.npz'))
    t = time.time()
    tab, cube = psfrec.compute_psf(LBDA35, 1.0, 0.7, 25., verbose=False)
    print('compute_psf wall %.3f s' % (time.time() - t))
    stat('compute_psf cube', cube, go['conv'])
    stat('compute_psf fwhm', tab['fwhm'][:, 0], go['fwhm'])
    stat('compute_psf n', tab['n'], go['n'])
    g4, o4 = np.load(os.path.join(GOLD, 'ref_config4_sample.npz')), np.load(os.path.join(GOLD, 'oracle_config4_sample.npz'))
    pick = g4['pick']
    fit, cube = psfrec.compute_psf_batch(g4['lbda'], g4['seeing'][pick], g4['GL'][pick], g4['L0'][pick],
                                         h=np.stack([g4['h0'][pick], g4['h1'][pick]], axis=1))
    stat('batch cfg4 cube', cube, o4['conv'])
    stat('batch cfg4 fwhm', fit[:, :, _lib.FIT_FWHM] * 0.2, o4['fwhm'])
    stat('batch cfg4 n', fit[:, :, _lib.FIT_N], o4['n'])
    o3 = np.load(os.path.join(GOLD, 'oracle_config3.npz'))
    tab3, cube3 = psfrec.compute_psf(o3['lbda'], 1.0, 0.7, 25., npsflin=3, three_lgs_mode=True, verbose=False)
    stat('compute_psf cfg3 cube', cube3, o3['conv'])
    stat('compute_psf cfg3 fwhm', tab3['fwhm'][:, 0], o3['fwhm'])


def s_misc():
    go = np.load(os.path.join(GOLD, 'oracle_config1.npz'))
    pol = psfrec.fit_psf_with_polynom(LBDA35, go['fwhm'], go['n'], output=1)
    stat('polyfit fwhm', pol['fwhm_pol'], go['ref_fwhm_pol'])
    stat('polyfit beta', pol['beta_pol'], go['ref_beta_pol'])
    ctx = psfrec.get_context()
    cubes = np.ascontiguousarray(np.stack([go['conv'], go['conv'][::-1] * 1.0, go['conv'] * 0.5]))
    mean = np.empty_like(go['conv'])
    fit = np.empty((35, _lib.FIT_NPAR))
    ctx.mean_refit(3, 35, cubes, mean, fit)
    stat('mean of cubes', mean, cubes.mean(axis=0))


def s_throughput():
    import torch
    rng = np.random.default_rng(12345)
    nd = 256
    seeing, GL, L0 = rng.uniform(0.4, 2.0, nd), rng.uniform(0.3, 0.95, nd), rng.uniform(9, 29, nd)
    h = np.stack([rng.uniform(50, 500, nd), rng.uniform(5000, 15000, nd)], axis=1)
    for rep in range(3):
        torch.cuda.synchronize()
        t = time.time()
        fit, cube = psfrec.compute_psf_batch(LBDA35, seeing, GL, L0, h=h, max_planes=64)
        torch.cuda.synchronize()
        dt = time.time() - t
        ctx = psfrec.get_context()
        ms, n, psfs = ctx.last_hot_timing()
        print('batch %d draws x 35: wall %.3f s -> %.0f PSF/s; hot kernel %.2f ms over %d launches (%d PSFs) -> %.0f PSF/s'
              % (nd, dt, nd * 35 / dt, ms, n, psfs, psfs / ms * 1e3), flush=True)
    print('fit iters range', fit[:, :, _lib.FIT_ITER].min(), fit[:, :, _lib.FIT_ITER].max(),
          'fwhm range', (fit[:, :, _lib.FIT_FWHM] * 0.2).min(), (fit[:, :, _lib.FIT_FWHM] * 0.2).max())
    print('kernel launches so far', ctx.kernel_launches())


if __name__ == '__main__':
    print('lib', _lib.LIB_PATH, os.path.exists(_lib.LIB_PATH))
    for fn in (s_otf, s_psd, s_structure, s_full_psf, s_psf_muse, s_convolve_fit, s_compute_psf, s_misc, s_throughput):
        section(fn)
