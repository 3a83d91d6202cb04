"""Time the Moffat fitter alone on the cubes of one chunk of the config-4 sweep (tuning aid):
    python tools/fit_bench.py [ndraw]      -> ms per call, mean iterations, largest FWHM / beta deviation from the
                                               default build's own result file if present"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from muse_psfr_b200 import _lib, psfrec

nd = int(sys.argv[1]) if len(sys.argv) > 1 else 128
psfrec.set_device(0)
seeing, GL, L0, h = bench.draws_for(12345, nd)
fit0, cube = psfrec.compute_psf_batch(bench.LBDA, seeing, GL, L0, h=h)
ctx = psfrec.get_context()
d_cube = torch.from_numpy(cube).cuda()
d_fit = torch.empty((nd * 35, _lib.FIT_NPAR), dtype=torch.float64, device='cuda')
for _ in range(3):
    ctx.moffat_fit(nd * 35, 40, 40, d_cube, d_fit)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ctx.moffat_fit(nd * 35, 40, 40, d_cube, d_fit)
e1.record()
torch.cuda.synchronize()
fit = d_fit.cpu().numpy()
ref_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gpurun_out', 'fit_ref.npy')
dev = ''
if os.path.exists(ref_path) and os.environ.get('PSFR_LIB_TAG'):
    ref = np.load(ref_path)
    dev = ' max rel dev fwhm %.2e beta %.2e' % (np.abs(fit[:, 5] / ref[:, 5] - 1).max(), np.abs(fit[:, 4] / ref[:, 4] - 1).max())
elif not os.environ.get('PSFR_LIB_TAG'):
    os.makedirs(os.path.dirname(ref_path), exist_ok=True)
    np.save(ref_path, fit)
print('fit: %.3f ms per %d images, iterations mean %.2f max %d, not converged %d%s' % (
    e0.elapsed_time(e1) / 10, nd * 35, np.abs(fit[:, 7]).mean(), np.abs(fit[:, 7]).max(), (fit[:, 7] < 0).sum(), dev))
