"""Tuning aid: error of the graded-precision row kernel against its all-FP64 evaluation and
against the oracle, for a few (f32_rows, grade) thresholds.  Run on a GPU box:
    python tools/diag_grade.py
Prints, per threshold pair, the largest difference relative to the PSF peak and the largest
pointwise relative difference over pixels above 1e-6 of the peak."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import psfr_oracle as orc
from muse_psfr_b200 import psfrec, _lib

psfrec.set_device(0)
ctx = psfrec.get_context()
lam = np.array([490., 560., 640., 780., 930.])
cases = [(1.68, 19.9, 0.36), (1.57, 28.0, 0.88), (0.91, 16.9, 0.71), (1.0, 25.0, 0.7), (0.55, 21.8, 0.91),
         (1.92, 24.7, 0.78), (1.11, 24.2, 0.57)]


def run(psd, g, f):
    ctx.set_option(_lib.OPT_EXP_GRADE, g)
    ctx.set_option(_lib.OPT_F32_ROWS, f)
    return psfrec.psf_muse(psd, lam)


def errs(a, b):
    peak = np.abs(b).max(axis=(1, 2), keepdims=True)
    sig = np.abs(b) > 1e-6 * peak
    return float((np.abs(a - b) / peak).max()), float((np.abs(a - b)[sig] / np.abs(b)[sig]).max())


pairs = [(25, 30), (20, 25), (18, 22), (15, 20), (12, 16)]
worst = {p: [0.0, 0.0, 0.0] for p in pairs}
for seeing, L0, GL in cases:
    psd = orc.simul_psd_wfm([GL, 1 - GL], (100, 10000), seeing, L0)
    ref = orc.psf_muse(psd, lam)
    full = run(psd, 1e30, 1e30)
    e_full = errs(full, ref)
    print('seeing %.2f L0 %.1f GL %.2f: all-FP64 vs oracle peak-rel %.1e pointwise %.1e' % (seeing, L0, GL, *e_full))
    for p in pairs:
        out = run(psd, *p)
        e1, e2 = errs(out, full)
        e3 = errs(out, ref)[1]
        w = worst[p]
        w[0], w[1], w[2] = max(w[0], e1), max(w[1], e2), max(w[2], e3)
for p in pairs:
    print('grade %g f32_rows %g: vs all-FP64 peak-rel %.1e pointwise(>1e-6 peak) %.1e | vs oracle pointwise %.1e'
          % (p[0], p[1], *worst[p]))
ctx.set_option(_lib.OPT_EXP_GRADE, 20.0)
ctx.set_option(_lib.OPT_F32_ROWS, 25.0)
