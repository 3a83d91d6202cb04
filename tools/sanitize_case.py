"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): one batch of 3 draws x 4
wavelengths with 9 directions on the 1280 grid, one 2560-grid psf_muse, and the SPARTA shell."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muse_psfr_b200 import psfrec  # noqa: E402

lam = np.array([490., 640., 800., 930.])
fit, cube = psfrec.compute_psf_batch(lam, [0.6, 1.0, 1.7], [0.7, 0.5, 0.4], [25., 15., 20.], npsflin=3, max_planes=18)
print('batch ok', np.isfinite(cube).all(), fit[:, :, 5].round(3).tolist())
psd = psfrec.simul_psd_wfm([0.7, 0.3], (100, 10000), 1.0, 25., dim=2560, verbose=False)
print('psf_muse 2560', psfrec.psf_muse(psd[0], lam[:2]).sum(axis=(1, 2)))
res = psfrec.compute_psf_from_sparta(psfrec.create_sparta_table(nlines=2, bad_l0=True), lmin=500, lmax=900, nl=3,
                                     verbose=False)
print('sparta ok', [h.name for h in res])
psfrec.release_contexts()
