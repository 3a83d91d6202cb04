"""Shell around the hot path: SPARTA table I/O, row rejection, CLI (reference:
psfrec.py:981-1141, cli.py) - CPU tests of the host logic (no GPU call is made), mirroring
the reference's own tests where they do not need a reconstruction."""
import io
import logging
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

import psfr_oracle as orc
from muse_psfr_b200 import _fits, psfrec
from muse_psfr_b200.cli import main


def test_fits_roundtrip(tmp_path):
    t = _fits.table_to_hdu({'lbda': np.array([500., 700.]), 'fwhm': np.ones((2, 2)) * 0.5,
                            'row_idx': np.array([1, 2])}, meta={'SEEING': 1.25, 'NOTE': "it's"}, name='fit_rows')
    img = _fits.ImageHDU(np.arange(2 * 3 * 4, dtype=float).reshape(2, 3, 4), name='PSF_MEAN')
    path = tmp_path / 'x.fits'
    _fits.HDUList([_fits.PrimaryHDU(), t, img]).writeto(path)
    assert os.path.getsize(path) % 2880 == 0
    with pytest.raises(OSError):
        _fits.HDUList([_fits.PrimaryHDU()]).writeto(path)               # exists, no overwrite
    with _fits.open(path) as r:
        assert [h.name for h in r] == ['PRIMARY', 'FIT_ROWS', 'PSF_MEAN']
        assert len(r) == 3 and 'FIT_ROWS' in r
        d = r['FIT_ROWS'].data
        assert d['fwhm'].shape == (2, 2) and d['row_idx'].dtype == np.int64
        assert_allclose(d['lbda'], [500., 700.])
        assert r['FIT_ROWS'].header['SEEING'] == 1.25 and r['FIT_ROWS'].header['NOTE'] == "it's"
        assert np.array_equal(r['PSF_MEAN'].data, img.data)
        with pytest.raises(KeyError):
            r['NOPE']
    # header-only access, also from a file object and from bytes
    assert _fits.getheader(path, 1)['EXTNAME'] == 'FIT_ROWS'
    raw = path.read_bytes()
    assert _fits.open(io.BytesIO(raw))[2].data.shape == (2, 3, 4)
    assert _fits.open(raw, only={'PSF_MEAN'})[1].data is None


def test_fits_reads_foreign_table_types():
    """Big-endian float32 / int32 / string columns and HIERARCH keywords, as ESO raw files carry."""
    cards = [_fits._card('XTENSION', 'BINTABLE'), _fits._card('BITPIX', 8), _fits._card('NAXIS', 2),
             _fits._card('NAXIS1', 12), _fits._card('NAXIS2', 2), _fits._card('PCOUNT', 0), _fits._card('GCOUNT', 1),
             _fits._card('TFIELDS', 3), _fits._card('TTYPE1', 'LGS1_SEEING'), _fits._card('TFORM1', 'E'),
             _fits._card('TTYPE2', 'N'), _fits._card('TFORM2', 'J'), _fits._card('TTYPE3', 'TAG'),
             _fits._card('TFORM3', '4A'), _fits._card('EXTNAME', 'SPARTA_ATM_DATA'), 'END'.ljust(80)]
    hdr = ''.join(cards).encode()
    hdr += b' ' * ((-len(hdr)) % 2880)
    rows = np.zeros(2, dtype=[('a', '>f4'), ('b', '>i4'), ('c', 'S4')])
    rows['a'], rows['b'], rows['c'] = [0.5, 1.5], [7, 8], [b'ab', b'cd']
    data = rows.tobytes()
    data += b'\0' * ((-len(data)) % 2880)
    prim = ''.join([_fits._card('SIMPLE', True), _fits._card('BITPIX', 8), _fits._card('NAXIS', 0),
                    _fits._card('EXTEND', True), 'HIERARCH ESO OBS NAME = \'WFM-AO-N_1\''.ljust(80),
                    'HIERARCH ESO TEL AIRM START = 1.25'.ljust(80), 'END'.ljust(80)]).encode()
    prim += b' ' * ((-len(prim)) % 2880)
    hdul = _fits.open(prim + hdr + data)
    assert hdul[0].header['ESO OBS NAME'] == 'WFM-AO-N_1' and hdul[0].header['ESO TEL AIRM START'] == 1.25
    d = hdul['SPARTA_ATM_DATA'].data
    assert_allclose(d['LGS1_SEEING'], [0.5, 1.5]) and d['N'].tolist() == [7, 8] and d['TAG'][1] == b'cd'


def test_create_sparta_table(tmp_path):
    hdu = psfrec.create_sparta_table(nlines=3, seeing=0.9, L0=22, GL=0.6, bad_l0=True,
                                     outfile=str(tmp_path / 's.fits'))
    assert hdu.name == 'SPARTA_ATM_DATA' and len(hdu.data) == 3
    assert hdu.data.dtype.names[:3] == ('LGS1_SEEING', 'LGS1_TUR_GND', 'LGS1_L0')
    d = _fits.open(str(tmp_path / 's.fits'))['SPARTA_ATM_DATA'].data
    assert_allclose(d['LGS2_SEEING'], 0.9) and assert_allclose(d['LGS4_L0'], 150) is None
    assert_allclose(d['LGS3_L0'], 22) and assert_allclose(d['LGS1_TUR_GND'], 0.6) is None
    hdu.data[0]['LGS1_L0'] = 20           # the reference's tests edit rows in place (test_psfrec.py:49-55)
    assert hdu.data['LGS1_L0'][0] == 20


def test_select_sparta_rows_matches_oracle(caplog):
    rng = np.random.default_rng(20261018)                      # SURVEY config 2 recipe
    n = 30
    seeing = np.clip(rng.lognormal(np.log(0.8), 0.25, n), 0.4, 2.0)
    GL = np.clip(rng.normal(0.7, 0.1, n), 0.3, 0.95)
    L0 = np.clip(rng.normal(18, 5, n), 9, 29)
    vals = np.stack([seeing, GL, L0], axis=1)[:, None, :] * (1 + 0.03 * rng.standard_normal((n, 4, 3)))
    vals[[3, 11, 27], 3, 2] = 150
    vals[5, :, 2] = 1000                                        # a row with no valid laser
    for mean in (True, False):
        with caplog.at_level(logging.INFO, logger='muse_psfr.psfrec'):
            got = psfrec.select_sparta_rows(vals, mean_of_lgs=mean, verbose=True)
        ref = orc.select_sparta_rows(vals, mean_of_lgs=mean)
        assert len(got) == len(ref) and (len(got) == 29 if mean else len(got) > 100)
        for g, r in zip(got, ref):
            assert_allclose(g[:3], r[:3], rtol=0, atol=0)
            assert tuple(g[3:]) == tuple(r[3:])
    msgs = [r.message for r in caplog.records]
    assert '4/30 : Using only 3 values out of 4 after outliers rejection' in msgs
    assert '6/30 : No valid values, skipping this row' in msgs


def test_no_valid_rows_returns_none(tmp_path, caplog):
    """test_bad_l0_invalid (test_psfrec.py:93-100): None + the two log records, before any GPU work."""
    testfile = str(tmp_path / 'sparta.fits')
    psfrec.create_sparta_table(outfile=testfile, L0=1000)
    with caplog.at_level(logging.INFO, logger='muse_psfr'):
        assert psfrec.compute_psf_from_sparta(testfile) is None
    assert caplog.records[0].message == 'Processing SPARTA table with 1 values, njobs=1 ...'
    assert caplog.records[1].message == '1/1 : No valid values, skipping this row'
    assert caplog.records[2].message == 'No valid values'


def test_script_argument_errors():
    """test_script (test_psfrec.py:103-111): the three SystemExit paths need no reconstruction."""
    with pytest.raises(SystemExit, match='no input file provided'):
        main([])
    with pytest.raises(SystemExit, match='--values must contain a list.*'):
        main(['--values', '0.1,0.2'])
    with pytest.raises(SystemExit, match='No results'):
        main(['--values', '1,0.7,1000'])


def test_package_exports():
    import muse_psfr_b200 as pkg
    for name in ('compute_psf_from_sparta', 'compute_psf', 'create_sparta_table', 'fit_psf_with_polynom',
                 'reconstruct_psf'):
        assert callable(getattr(pkg, name))
    assert pkg.__version__


def test_copied_extension_travels_verbatim(tmp_path):
    """the SPARTA extension copied into the result keeps header cards and column types this
    reader does not model (TUNIT, TNULL, logical columns): it is re-emitted byte for byte."""
    import io
    from muse_psfr_b200 import _fits
    cards = [_fits._card('XTENSION', 'BINTABLE'), _fits._card('BITPIX', 8), _fits._card('NAXIS', 2),
             _fits._card('NAXIS1', 9), _fits._card('NAXIS2', 2), _fits._card('PCOUNT', 0), _fits._card('GCOUNT', 1),
             _fits._card('TFIELDS', 2), _fits._card('TTYPE1', 'LGS1_SEEING'), _fits._card('TFORM1', 'D'),
             _fits._card('TUNIT1', 'arcsec'), _fits._card('TTYPE2', 'FLAG'), _fits._card('TFORM2', 'L'),
             _fits._card('EXTNAME', 'SPARTA_ATM_DATA'), 'END'.ljust(80)]
    hdr = ''.join(cards).encode('ascii')
    hdr += b' ' * _fits._pad(len(hdr))
    payload = np.array([1.0, 0.8], dtype='>f8').tobytes()
    rows = payload[:8] + b'T' + payload[8:] + b'F'
    ext = hdr + rows + b'\0' * _fits._pad(len(rows))
    primary = _fits.HDUList([_fits.PrimaryHDU()]).tobytes()
    hdul = _fits.open(io.BytesIO(primary + ext))
    assert list(hdul['SPARTA_ATM_DATA'].data['LGS1_SEEING']) == [1.0, 0.8]
    out = _fits.HDUList([_fits.PrimaryHDU(), hdul['SPARTA_ATM_DATA'].copy()]).tobytes()
    assert out[len(primary):] == ext
    again = _fits.open(io.BytesIO(out))
    assert again['SPARTA_ATM_DATA'].header.get('EXTNAME') == 'SPARTA_ATM_DATA'
