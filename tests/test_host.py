"""CPU tests of the host side: the C-ABI library loads and exports every symbol declared in
include/psfr.h, there is no CPU fallback, and the host scalar bookkeeping matches the
reference's own expressions (through the oracle)."""
import ctypes
import os
import re

import numpy as np
import pytest
from numpy.testing import assert_allclose

import psfr_oracle as orc
from muse_psfr_b200 import _lib, psfrec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built_lib():
    from muse_psfr_b200.build import build
    return build()


def test_header_and_binding_agree():
    with open(os.path.join(ROOT, 'include', 'psfr.h')) as f:
        hdr = f.read()
    declared = set(re.findall(r'PSFR_API\s+[\w\s\*]+?\b(psfr_\w+)\s*\(', hdr))
    assert declared == set(_lib.exported_symbols())
    # record layouts
    for name in ('R0', 'L0', 'FITC', 'ALPHA_TT', 'NLAYERS', 'LAYER0', 'NPAR'):
        m = re.search(r'PSFR_DRAW_%s = (\d+)' % name, hdr)
        assert int(m.group(1)) == getattr(_lib, 'DRAW_' + name)
    for name in ('CPHI', 'H', 'WX', 'WY', 'NPAR'):
        m = re.search(r'PSFR_LAYER_%s = (\d+)' % name, hdr)
        assert int(m.group(1)) == getattr(_lib, 'LAYER_' + name)
    assert int(re.search(r'#define PSFR_MAX_LAYERS (\d+)', hdr).group(1)) == _lib.MAX_LAYERS
    assert _lib.layer_slot(_lib.MAX_LAYERS - 1, _lib.LAYER_WY) < _lib.DRAW_NPAR
    for name in ('PEAK', 'Y0', 'X0', 'ALPHA', 'N', 'FWHM', 'CHISQ', 'ITER', 'ERR_FWHM', 'FLUX', 'ERR_FLUX', 'NPAR'):
        m = re.search(r'PSFR_FIT_%s = (\d+)' % name, hdr)
        assert int(m.group(1)) == getattr(_lib, 'FIT_' + name)
    for name, key in _lib.INFO_KEYS.items():
        assert int(re.search(r'PSFR_INFO_%s = (\d+)' % name.upper(), hdr).group(1)) == key
    # option keys of psfr_set_option
    for name in ('EXP_CUT', 'EXP_GRADE', 'F32_ROWS', 'ROW_KERNEL'):
        m = re.search(r'PSFR_OPT_%s = (\d+)' % name, hdr)
        assert int(m.group(1)) == getattr(_lib, 'OPT_' + name)


def test_library_exports_every_symbol(built_lib):
    handle = ctypes.CDLL(built_lib)
    for sym in _lib.exported_symbols():
        assert hasattr(handle, sym), sym
    assert _lib.load().psfr_version() >= 100


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present: the fallback check is for CPU-only boxes')
    with pytest.raises(_lib.PsfrError) as exc:
        _lib.Context(device=0)
    assert 'no CUDA device' in str(exc.value) or 'CUDA' in str(exc.value)
    with pytest.raises(_lib.PsfrError):
        psfrec.compute_psf(np.array([500.]), 1.0, 0.7, 25., verbose=False)


def test_unsupported_dim_is_loud(built_lib):
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.psfr_create(0, 640, 4, 4, ctypes.byref(h))
    assert rc == _lib.E_UNSUPPORTED and not h.value
    assert b'dim=640' in lib.psfr_last_error(None)
    with pytest.raises(NotImplementedError):
        psfrec.simul_psd_wfm([0.7, 0.3], (100, 10000), 1.0, 25., dim=640)


def test_draw_record_matches_reference_scalars():
    rec = psfrec.draw_record([0.7, 0.3], (100, 10000), 1.0, 25.)
    r0 = orc.seeing2r01(1.0, 0.5, 0.)
    assert rec[_lib.DRAW_R0] == r0
    assert rec[_lib.DRAW_NLAYERS] == 2
    wind = orc.wind_speed_for((100, 10000))
    assert rec[_lib.layer_slot(0, _lib.LAYER_WX)] == wind[0] * np.cos(orc.WIND_DIR[0])        # 12 m/s: integer altitudes
    assert rec[_lib.layer_slot(1, _lib.LAYER_WY)] == wind[1] * np.sin(orc.WIND_DIR[1])
    recf = psfrec.draw_record([0.7, 0.3], (100., 10000.), 1.0, 25.)
    assert recf[_lib.layer_slot(0, _lib.LAYER_WX)] == 12.5 * np.cos(orc.WIND_DIR[0])
    cn2 = np.array([0.7, 0.3])
    cn2 /= cn2.sum()
    assert rec[_lib.layer_slot(1, _lib.LAYER_CPHI)] == 0.0229 * (cn2[1] ** (-3 / 5) * r0) ** (-5 / 3)
    with pytest.raises(ValueError):
        psfrec.draw_record([0.5, 0.3, 0.2], (100, 5000, 10000), 1.0, 25.)   # reference: ValueError too
    # extension: explicit wind directions lift the limit, the zenith angle enters r0
    rec3 = psfrec.draw_record([0.5, 0.3, 0.2], (100., 5000., 10000.), 1.0, 25., zenith=30., wind_dir=[0.1, 0.2, 0.3])
    assert rec3[_lib.DRAW_NLAYERS] == 3 and rec3[_lib.DRAW_R0] == orc.seeing2r01(1.0, 0.5, 30.)
    assert rec3[_lib.layer_slot(2, _lib.LAYER_WY)] == 12.5 * np.sin(0.3)
    with pytest.raises(ValueError):
        psfrec.draw_record(np.ones(9), np.arange(9.), 1.0, 25., wind_dir=np.zeros(9))


def test_vectorised_records_agree_with_scalar():
    rng = np.random.default_rng(3)
    n = 50
    s, g, l0 = rng.uniform(.4, 2, n), rng.uniform(.3, .95, n), rng.uniform(9, 29, n)
    h = np.stack([rng.uniform(50, 500, n), rng.uniform(5000, 15000, n)], 1)
    recs = psfrec.draw_records(s, g, l0, h)
    for i in range(n):
        r = psfrec.draw_record([g[i], 1 - g[i]], h[i], s[i], l0[i], 0.,
                               alpha_tt=psfrec.tiptilt_alpha(s[i], g[i], l0[i]))
        assert_allclose(recs[i], r, rtol=2e-15, atol=0)


def test_tiptilt_alpha_and_tables():
    assert psfrec.tiptilt_alpha(1.0, 0.7, 25.) == orc.tiptilt_alpha(1.0, 0.7, 25.)
    assert_allclose(psfrec._coeff_hl(25.), 0.28365585, rtol=1e-7)     # SURVEY a10 probe values
    assert_allclose(psfrec._coeff_hl(8.), 0.03942410, rtol=1e-6)
    f, fx, fy = psfrec.ao_frequency_tables()
    fo, _, fxo, fyo = orc.ao_frequency_tables()
    assert np.array_equal(f, fo) and np.array_equal(fx, fxo) and np.array_equal(fy, fyo)
    assert np.array_equal(psfrec.direction_perf(3), orc.direction_perf(3))
    assert np.array_equal(psfrec.pupil_mask(320, 640, oc=0.14), orc.pupil_mask(320, 640, 0.14))
    fw, be, _, _ = psfrec.muse_intrinsic_psf(np.linspace(490, 930, 35))
    fwo, beo = orc.muse_intrinsic_psf(np.linspace(490, 930, 35))
    assert np.array_equal(fw, fwo) and np.array_equal(be, beo)


def test_fit_table_behaves_like_a_table():
    fit = np.zeros((3, _lib.FIT_NPAR))
    fit[:, _lib.FIT_FWHM] = [4., 3.5, 3.]
    fit[:, _lib.FIT_N] = [2.7, 2.5, 2.2]
    tab = psfrec._table_from_fit([500., 700., 900.], fit)
    assert len(tab) == 3
    assert tab.colnames[:6] == ['lbda', 'center', 'flux', 'fwhm', 'n', 'peak']
    assert_allclose(tab['fwhm'][:, 0], [0.8, 0.7, 0.6])
    tab['SEEING'] = 1.0
    assert tab[1]['SEEING'] == 1.0 and tab[1]['n'] == 2.5
    both = psfrec.FitTable.vstack([tab, tab])
    assert len(both) == 6


def test_caller_buffers_are_validated_at_the_binding():
    """the C ABI trusts pointers; the binding must reject buffers of the wrong dtype,
    size, layout or device before the library writes through them."""
    import torch
    ok = np.empty((2, 3, 40, 40))
    assert _lib.checked_ptr(ok, 2 * 3 * 1600, 0, 'out_cube', True) == ok.ctypes.data
    assert _lib.checked_ptr(None, 10, 0, 'out') is None
    with pytest.raises(ValueError, match='float64'):
        _lib.checked_ptr(ok.astype(np.float32), 10, 0, 'out_cube', True)
    with pytest.raises(ValueError, match='elements'):
        _lib.checked_ptr(ok, ok.size + 1, 0, 'out_cube', True)
    with pytest.raises(ValueError, match='contiguous'):
        _lib.checked_ptr(ok[:, :, ::2], 10, 0, 'out_cube', True)
    ro = ok.copy()
    ro.flags.writeable = False
    with pytest.raises(ValueError, match='read-only'):
        _lib.checked_ptr(ro, 10, 0, 'out_cube', True)
    t = torch.empty(100, dtype=torch.float64)
    assert _lib.checked_ptr(t, 100, 0, 'out_fit', True) == t.data_ptr()
    with pytest.raises(ValueError, match='float64'):
        _lib.checked_ptr(t.float(), 100, 0, 'out_fit', True)
    with pytest.raises(ValueError, match='elements'):
        _lib.checked_ptr(t, 101, 0, 'out_fit', True)
    with pytest.raises(ValueError, match='contiguous'):
        _lib.checked_ptr(t[::2], 10, 0, 'out_fit', True)
    with pytest.raises(TypeError):
        _lib.checked_ptr([1.0, 2.0], 2, 0, 'out_fit', True)


def test_empty_and_ragged_batches(built_lib):
    """No draws / no wavelengths give empty results of the right shape without touching the device;
    parameter arrays of different lengths are refused before anything is computed."""
    lam = np.linspace(490, 930, 35)
    fit, cube = psfrec.compute_psf_batch(lam, [], [], [])
    assert fit.shape == (0, 35, _lib.FIT_NPAR) and cube.shape == (0, 35, 40, 40)
    fit, cube = psfrec.compute_psf_batch([], [1.0], [0.7], [25.], want_cube=False)
    assert fit.shape == (1, 0, _lib.FIT_NPAR) and cube is None
    with pytest.raises(ValueError, match='one entry per draw'):
        psfrec.compute_psf_batch(lam, [1.0, 0.8], [0.7], [25., 20.])
