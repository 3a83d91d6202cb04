"""GPU tests of the shell (compute_psf_from_sparta + CLI): the reference's own end-to-end
tests (muse_psfr/test_psfrec.py) restated against this backend, plus a parity check of a
multi-row table against the oracle driven the way the reference drives compute_psf."""
import logging
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

import psfr_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def pkg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import muse_psfr_b200 as mod
    from muse_psfr_b200 import psfrec
    psfrec.set_device(0)
    yield mod
    psfrec.release_contexts()


def test_reconstruction(pkg):                                   # test_psfrec.py:17-30
    from muse_psfr_b200 import _fits
    tbl = pkg.create_sparta_table()
    hdul = _fits.HDUList([_fits.PrimaryHDU(), tbl])
    res = pkg.compute_psf_from_sparta(hdul, npsflin=3, lmin=490, lmax=541.76, nl=5)
    assert len(res) == 5
    fit = res['FIT_ROWS'].data
    assert_allclose(fit['L0'], 25)
    assert_allclose(fit['center'], 20, atol=1e-6)
    assert_allclose(fit[1]['lbda'], 502.9, atol=1e-1)
    assert_allclose(fit[1]['fwhm'], 0.85, atol=1e-2)


def test_fit_poly(pkg):                                         # test_psfrec.py:33-44
    res = pkg.compute_psf_from_sparta(pkg.create_sparta_table(), lmin=500, lmax=900, nl=9)
    fit = res['FIT_ROWS'].data
    res = pkg.fit_psf_with_polynom(fit['lbda'], fit['fwhm'][:, 0], fit['n'], deg=(5, 5), output=1)
    assert_allclose(res['fwhm_pol'][0], 0.65, atol=1e-2)
    assert_allclose(res['beta_pol'][0], 0.78, atol=1e-2)
    assert_allclose(res['beta_fit'][8], fit[1]['n'], atol=1e-2)
    assert_allclose(res['fwhm_fit'][8], fit[1]['fwhm'], atol=1e-2)


def test_reconstruction2(pkg):                                  # test_psfrec.py:47-69
    tbl = pkg.create_sparta_table()
    tbl.data[0]['LGS1_L0'] = 20
    tbl.data[0]['LGS1_SEEING'] = 0.8
    tbl.data[0]['LGS1_TUR_GND'] = 0.5
    tbl.data[0]['LGS3_L0'] = 100
    res = pkg.compute_psf_from_sparta(tbl, npsflin=3, lmin=500, lmax=700, nl=3, mean_of_lgs=False)
    assert len(res) == 5
    fit = res['FIT_ROWS'].data
    assert_allclose(fit[fit['lgs_idx'] == 1]['L0'], 20)
    assert_allclose(fit[fit['lgs_idx'] != 1]['L0'], 25)
    assert_allclose(fit['center'], 20, atol=1e-6)
    assert_allclose(fit[fit['lbda'] == 500]['fwhm'][:, 0], [0.79, 0.86, 0.86], atol=1e-2)


def test_bad_l0(pkg, tmp_path, caplog):                         # test_psfrec.py:72-90
    testfile = os.path.join(str(tmp_path), 'sparta.fits')
    pkg.create_sparta_table(outfile=testfile, bad_l0=True)
    with caplog.at_level(logging.INFO, logger='muse_psfr'):
        res = pkg.compute_psf_from_sparta(testfile, lmin=490, lmax=541.76, nl=5)
    assert caplog.records[1].message == '1/1 : Using only 3 values out of 4 after outliers rejection'
    assert caplog.records[3].message == 'Using three lasers mode'
    assert len(res) == 5
    fit = res['FIT_ROWS'].data
    assert_allclose(fit['L0'], 25)
    assert_allclose(fit['center'], 20, atol=1e-6)
    assert_allclose(fit[1]['lbda'], 502.9, atol=1e-1)
    assert_allclose(fit[1]['fwhm'], 0.86, atol=1e-2)


def test_script(pkg, tmp_path, caplog):                         # test_psfrec.py:103-149
    from muse_psfr_b200.cli import main
    logfile = os.path.join(str(tmp_path), 'muse-psfr2.log')
    with caplog.at_level(logging.INFO, logger='muse_psfr'):
        main(['--no-color', '--values', '1,0.7,25', '--logfile', logfile])
    with open(logfile) as f:
        lines = f.read().splitlines()
    assert lines[2:] == [
        '--------------------------------------------------------------------',
        'Sparta Seeing: 1.00 arcsec GL: 0.70 L0:25.00 m',
        'LBDA 5000 7000 9000',
        'FWHM 0.85 0.73 0.62',
        'BETA 2.73 2.55 2.23',
        '--------------------------------------------------------------------'
    ]
    records = [r for r in caplog.records if r.levelname != 'DEBUG']
    assert records[6].message == 'LBDA 5000 7000 9000'
    assert records[7].message == 'FWHM 0.85 0.73 0.62'
    assert records[8].message == 'BETA 2.73 2.55 2.23'


def test_script_with_file(pkg, tmp_path):                       # test_psfrec.py:152-170
    from muse_psfr_b200 import _fits
    from muse_psfr_b200.cli import main
    testfile = os.path.join(str(tmp_path), 'sparta.fits')
    pkg.create_sparta_table(outfile=testfile)
    logfile = os.path.join(str(tmp_path), 'muse_psfr.log')
    outfile = os.path.join(str(tmp_path), 'out.fits')
    main([testfile, '--no-color', '--logfile', logfile, '--outfile', outfile])
    with _fits.open(outfile) as hdul:
        assert [hdu.name for hdu in hdul] == ['PRIMARY', 'SPARTA_ATM_DATA', 'FIT_ROWS', 'FIT_MEAN', 'PSF_MEAN']
        assert hdul['PSF_MEAN'].data.shape == (3, 40, 40)
        assert hdul['FIT_ROWS'].data.dtype.names == (
            'lbda', 'center', 'flux', 'fwhm', 'n', 'peak', 'err_center', 'err_flux', 'err_fwhm', 'err_n',
            'err_peak', 'converged', 'SEEING', 'GL', 'L0', 'row_idx', 'lgs_idx')
    with open(logfile) as f:
        lines = f.read().splitlines()
    assert lines[2:] == [
        'OB None None Airmass 0.00-0.00',
        '--------------------------------------------------------------------',
        'Sparta Seeing: 1.00 arcsec GL: 0.70 L0:25.00 m',
        'LBDA 5000 7000 9000',
        'FWHM 0.85 0.73 0.62',
        'BETA 2.73 2.55 2.23',
        '--------------------------------------------------------------------'
    ]


def test_sparta_table_parity_with_oracle(pkg):
    """BASELINE config 2 in small: a jittered multi-row table (one row in 3-LGS mode, one row
    rejected), every row + the time-mean PSF and its refit against the oracle driven like the
    reference (row selection -> compute_psf per row -> mean -> fit)."""
    from muse_psfr_b200 import _fits
    rng = np.random.default_rng(20261018)
    n = 5
    seeing = np.clip(rng.lognormal(np.log(0.8), 0.25, n), 0.4, 2.0)
    GL = np.clip(rng.normal(0.7, 0.1, n), 0.3, 0.95)
    L0 = np.clip(rng.normal(18, 5, n), 9, 29)
    vals = np.stack([seeing, GL, L0], axis=1)[:, None, :] * (1 + 0.03 * rng.standard_normal((n, 4, 3)))
    vals[1, 3, 2] = 150          # 3-LGS mode
    vals[3, :, 2] = 500          # rejected
    cols = {'LGS%d_%s' % (k + 1, c): vals[:, k, j] for k in range(4) for j, c in enumerate(('SEEING', 'TUR_GND', 'L0'))}
    hdu = _fits.table_to_hdu(cols, name='SPARTA_ATM_DATA')
    lam = np.array([490., 700., 930.])
    res = pkg.compute_psf_from_sparta(hdu, lbda=lam, verbose=False)
    jobs = orc.select_sparta_rows(vals, mean_of_lgs=True)
    assert len(jobs) == 4
    fit = res['FIT_ROWS'].data
    cubes = []
    for j, (s, g, l0, three, irow, lgs) in enumerate(jobs):
        ref, cube = orc.compute_psf(lam, s, g, l0, three_lgs_mode=three)
        cubes.append(cube)
        rows = fit[fit['row_idx'] == j + 1]
        assert_allclose(rows['fwhm'][:, 0], ref['fwhm'], rtol=1e-5)
        assert_allclose(rows['n'], ref['n'], rtol=1e-5)
        assert_allclose(rows['SEEING'], s) and assert_allclose(rows['lgs_idx'], -1) is None
    mean = np.mean(cubes, axis=0)
    got_mean = res['PSF_MEAN'].data
    assert np.abs(got_mean - mean).max() / mean.max() < 1e-9
    ref_mean = orc.fit_psf_cube(lam, mean)
    fm = res['FIT_MEAN'].data
    assert_allclose(fm['fwhm'][:, 0], ref_mean['fwhm'], rtol=1e-5)
    assert_allclose(fm['n'], ref_mean['n'], rtol=1e-5)
    med = np.median([[j[0], j[1], j[2]] for j in jobs], axis=0)
    hdr = res['FIT_MEAN'].header
    assert_allclose([hdr['SEEING'], hdr['GL'], hdr['L0']], med, rtol=1e-12)
