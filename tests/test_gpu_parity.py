"""GPU parity tests: the CUDA path (through the C ABI and the reference-shaped Python
functions) against the CPU oracle on the same inputs and against the committed fixtures
generated from the reference's own code.

Tolerances (BASELINE.json): PSF images relative 1e-9 in FP64; fitted FWHM / beta relative 1e-5.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import psfr_oracle as orc

pytestmark = pytest.mark.gpu

LBDA35 = np.linspace(490, 930, 35)
PSF_RTOL = 1e-9      # north_star: PSF images to relative 1e-9
FIT_RTOL = 1e-5      # north_star: FWHM / beta to relative 1e-5


@pytest.fixture(scope='module')
def psfrec():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from muse_psfr_b200 import psfrec as mod
    mod.set_device(0)
    yield mod
    mod.release_contexts()


@pytest.fixture(scope='module')
def psd1():
    return orc.simul_psd_wfm([0.7, 0.3], (100, 10000), 1.0, 25.)


def rel_to_peak(got, ref):
    return np.abs(np.asarray(got) - np.asarray(ref)).max() / np.abs(ref).max()


def assert_image_close(got, ref, rtol=PSF_RTOL, floor=1e-6):
    """Relative 1e-9: every pixel above `floor` (1e-6) of the peak agrees to rtol pointwise, and
    the whole image agrees to rtol of the peak."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape
    assert np.isfinite(got).all()
    assert rel_to_peak(got, ref) < rtol
    sig = np.abs(ref) > floor * np.abs(ref).max()
    assert_allclose(got[sig], ref[sig], rtol=rtol, atol=0)


# ---------------------------------------------------------------- init constants
def test_telescope_otf(psfrec):
    t = psfrec.get_context().get_otf()
    ref = orc.telescope_otf(orc.pupil_mask(320, 640, 0.14), 1280)
    assert rel_to_peak(t[:641], ref[:641]) < 1e-13
    assert np.all(t[641] == 0)
    assert_allclose(t[640, 640] * 1280 ** 2, 1.0, rtol=1e-15)
    # exact zeros outside the pupil-autocorrelation support (disc of radius N/2)
    yy, xx = np.ogrid[:641, :1280]
    assert np.all(t[:641][np.hypot(yy - 640, xx - 640) > 641] == 0)


# ---------------------------------------------------------------- PSD synthesis (a1-a7)
@pytest.mark.parametrize('args,kw', [
    (([0.7, 0.3], (100, 10000), 1.0, 25.), dict(npsflin=1)),
    (([0.7, 0.3], (100, 10000), 1.0, 25.), dict(npsflin=3, three_lgs_mode=True)),
    (([0.55, 0.45], (150.5, 12000.), 0.63, 12.5), dict(npsflin=2)),
    (([1.0], (300.,), 1.4, 9.5), dict(npsflin=1)),
])
def test_simul_psd_wfm(psfrec, args, kw):
    got = psfrec.simul_psd_wfm(*args, verbose=False, **kw)
    ref = orc.simul_psd_wfm(*args, **kw)
    assert got.shape == ref.shape
    assert np.array_equal(got == 0, ref == 0)            # cut-off masks bit-exact
    assert_allclose(got, ref, rtol=1e-10, atol=0)
    assert_allclose(got.sum(axis=(1, 2)), ref.sum(axis=(1, 2)), rtol=1e-12)


def test_psd_golden_config3(psfrec, golden):
    g = golden('ref_config3')
    got = psfrec.simul_psd_wfm([0.7, 0.3], (100, 10000), 1.0, 25., npsflin=3, three_lgs_mode=True, verbose=False)
    assert_allclose(got[:, 600:680, 600:680], g['psd_aozone'], rtol=1e-10)
    assert_allclose(got.sum(axis=(1, 2)), g['psd_sum'], rtol=1e-12)


def test_three_layers_rejected(psfrec):
    with pytest.raises(ValueError):
        psfrec.simul_psd_wfm([0.5, 0.3, 0.2], (100, 5000, 10000), 1.0, 25.)


def test_more_layers_and_zenith(psfrec):
    """layers beyond the reference's two (explicit wind directions) and zenith != 0,
    against the oracle extended the same way (the reference's formulas are layer-count agnostic)."""
    cn2, h, wd = [0.5, 0.3, 0.15, 0.05], (80., 2500., 9000., 14000.), [0.628163, -0.326497, 1.2, -2.0]
    got = psfrec.simul_psd_wfm(cn2, h, 0.9, 18., zenith=30., npsflin=2, wind_dir=wd)
    ref = orc.simul_psd_wfm(cn2, h, 0.9, 18., zenith=30., npsflin=2, wind_dir=wd)
    assert np.array_equal(got == 0, ref == 0)
    assert_allclose(got, ref, rtol=1e-10, atol=0)
    lam = np.array([490., 700., 930.])
    tab, cube = psfrec.compute_psf(lam, 0.9, 0.6, 18., h=h[:3], Cn2=cn2[:3], wind_dir=wd[:3], zenith=30., verbose=False)
    rtab, rcube = orc.compute_psf(lam, 0.9, 0.6, 18., h=h[:3], Cn2=cn2[:3], wind_dir=wd[:3], zenith=30.)
    for i in range(lam.size):
        assert_image_close(cube[i], rcube[i])
    assert_allclose(tab['fwhm'][:, 0], rtab['fwhm'], rtol=FIT_RTOL)
    assert_allclose(tab['n'], rtab['n'], rtol=FIT_RTOL)
    # zenith alone on the reference's own two-layer profile
    tab0, cube0 = psfrec.compute_psf(lam[:1], 1.0, 0.7, 25., zenith=30., verbose=False)
    rtab0, rcube0 = orc.compute_psf(lam[:1], 1.0, 0.7, 25., zenith=30.)
    assert_image_close(cube0[0], rcube0[0])
    assert_allclose(tab0['fwhm'][:, 0], rtab0['fwhm'], rtol=FIT_RTOL)


# ---------------------------------------------------------------- BASELINE config 5: dim = 2560
@pytest.fixture(scope='module')
def psd5(psfrec):
    return psfrec.simul_psd_wfm([0.7, 0.3], (100, 10000), 1.0, 25., dim=2560, verbose=False)


def test_config5_psd_golden(psd5, golden):
    """simul_psd_wfm(dim=2560) against the reference's own output (reduced fixture)."""
    g = golden('ref_config5')
    a, n, c = psd5[0], 2560, 1280
    assert psd5.shape == (1, n, n)
    assert_allclose(a[c - 48:c + 48, c - 48:c + 48], g['psd_centre'], rtol=1e-10)
    assert_allclose(a[::16, ::16], g['psd_lattice'], rtol=1e-10)
    assert_allclose(a[[0, 1, c - 1, c, c + 1, n - 1]], g['psd_rows'], rtol=1e-10)
    assert_allclose(a.sum(), g['psd_sum'], rtol=1e-12)
    assert_allclose(a.max(), g['psd_max'], rtol=1e-12)


def test_config5_telescope_otf(psfrec):
    t = psfrec.get_context(dim=2560).get_otf()
    ref = orc.telescope_otf(orc.pupil_mask(640, 1280, 0.14), 2560)
    assert t.shape == (1282, 2560)
    assert rel_to_peak(t[:1281], ref[:1281]) < 1e-13
    assert np.all(t[1281] == 0)
    assert_allclose(t[1280, 1280] * 2560 ** 2, 1.0, rtol=1e-15)


def test_config5_psf_muse_golden(psfrec, psd5, golden):
    """psf_muse on the 2560 grid (two 1280-point warp transforms + radix-2 combine per line)
    against the reference's own output at the two ends of the 100-wavelength range."""
    g = golden('ref_config5')
    got = psfrec.psf_muse(psd5[0], g['lbda'])
    assert got.shape == g['psf_muse'].shape
    for k in range(got.shape[0]):
        assert_image_close(got[k], g['psf_muse'][k])


def test_config5_psd_to_psf_full_grid(psfrec, psd5):
    pup = orc.pupil_mask(640, 1280, 0.14)
    got = psfrec.psd_to_psf(psd5[0], pup, 8, 700e-9, samp=2)
    ref = orc.psd_to_psf(psd5[0], pup, 8, 700e-9)
    # two FP64 2560^2 transforms (pocketfft vs. ours) differ by ~4e-19 absolute = 1e-15 of the peak,
    # which is 1.2e-9 of the faintest pixels above 1e-6 of the peak: pointwise bar from 1e-5 up here
    assert rel_to_peak(got, ref) < 1e-13
    assert_image_close(got, ref, floor=1e-5)
    assert_allclose(got.sum(), 1.0, rtol=1e-12)


def test_config5_hundred_wavelengths(psfrec, psd5):
    """All 100 wavelengths of config 5 in one call; the pruned path equals the full-grid path
    resampled by the oracle's psf_muse tail at a few of them."""
    lam = np.linspace(490, 930, 100)
    cube = psfrec.psf_muse(psd5[0], lam)
    assert cube.shape == (100, 40, 40) and np.isfinite(cube).all()
    assert_allclose(cube.sum(axis=(1, 2)), 1.0, rtol=1e-12)
    ref = orc.psf_muse(psd5[0], lam[[0, 57]])
    assert_image_close(cube[0], ref[0])
    assert_image_close(cube[57], ref[1])


def test_config5_compute_psf(psfrec):
    """compute_psf(dim=2560): the fitted FWHM moves by < 1 % against the 1280 grid (the wider
    fitting PSD only adds far wings) and matches the oracle run at dim=2560."""
    lam = np.array([500., 700., 900.])
    tab, cube = psfrec.compute_psf(lam, 1.0, 0.7, 25., verbose=False, dim=2560)
    ref, ref_cube = orc.compute_psf(lam, 1.0, 0.7, 25., dim=2560)
    for k in range(3):
        assert_image_close(cube[k], ref_cube[k])
    assert_allclose(tab['fwhm'][:, 0], ref['fwhm'], rtol=FIT_RTOL)
    assert_allclose(tab['n'], ref['n'], rtol=FIT_RTOL)
    tab1, _ = psfrec.compute_psf(lam, 1.0, 0.7, 25., verbose=False)
    assert_allclose(tab['fwhm'][:, 0], tab1['fwhm'][:, 0], rtol=1e-2)


# ---------------------------------------------------------------- structure function / psd_to_psf (a8)
def test_structure_function(psfrec, psd1):
    ctx = psfrec.get_context()
    ctx.load_psd(np.ascontiguousarray(psd1[0]), 1)
    ctx.structure_function(1)
    d = ctx.get_structure_function(0)
    ref = orc.structure_function_unit(psd1[0])
    assert rel_to_peak(d[:641], ref.T[:641]) < 1e-13
    assert d[640, 640] == 0.0 and np.all(d[641] == 0)


@pytest.mark.parametrize('lb', [490e-9, 710e-9, 930e-9])
def test_psd_to_psf_full_grid(psfrec, psd1, golden, lb):
    pup = orc.pupil_mask(320, 640, 0.14)
    got = psfrec.psd_to_psf(psd1[0], pup, 8, lb, samp=2)
    ref = orc.psd_to_psf(psd1[0], pup, 8, lb)
    assert_image_close(got, ref)
    assert_allclose(got.sum(), 1.0, rtol=1e-12)
    g = golden('ref_config1')
    name = 'psf%d' % round(lb * 1e9)
    assert_allclose(got[640 - 48:640 + 48, 640 - 48:640 + 48], g[name + '_centre'], rtol=PSF_RTOL)
    assert rel_to_peak(got[::16, ::16], g[name + '_lattice']) < PSF_RTOL


def test_psd_to_psf_dead_branches(psfrec, psd1):
    pup = orc.pupil_mask(320, 640, 0.14)
    with pytest.raises(NotImplementedError):
        psfrec.psd_to_psf(psd1[0], pup, 8, 500e-9, samp=3)
    with pytest.raises(NotImplementedError):
        psfrec.psd_to_psf(psd1[0], pup, 8, 500e-9, samp=2, FoV=10.)
    with pytest.raises(NotImplementedError):
        psfrec.psd_to_psf(psd1[0][:640, :640], pup[:320, :320], 8, 500e-9)


# ---------------------------------------------------------------- psf_muse (a9)
def test_psf_muse_config1_golden(psfrec, psd1, golden):
    g = golden('ref_config1')
    got = psfrec.psf_muse(psd1[0], LBDA35)
    assert got.shape == (35, 40, 40)
    for i in range(35):
        assert_image_close(got[i], g['psf_muse'][i])
    assert_allclose(got.sum(axis=(1, 2)), 1.0, rtol=1e-13)
    assert_allclose(got[0].max(), 0.06878001975345309, rtol=1e-9)     # SURVEY 8c probes
    assert_allclose(got[34].max(), 0.1438347394519584, rtol=1e-9)


def test_psf_muse_nine_directions(psfrec, golden):
    g = golden('ref_config3')
    psd3 = orc.simul_psd_wfm([0.7, 0.3], (100, 10000), 1.0, 25., npsflin=3, three_lgs_mode=True)
    got = psfrec.psf_muse(psd3, g['lbda'])
    for i in range(len(g['lbda'])):
        assert_image_close(got[i], g['psf_muse'][i])


def test_psf_muse_wavelength_order_and_count(psfrec, psd1):
    """The row kernel works through the wavelengths in sorted order and stages their scalars in
    shared memory up to 64 of them: an unsorted list, and a list longer than 64, must give the
    same planes as one-at-a-time calls."""
    rng = np.random.default_rng(7)
    lam = rng.permutation(np.linspace(495, 925, 70))
    cube = psfrec.psf_muse(psd1[0], lam)
    assert cube.shape == (70, 40, 40)
    for k in (0, 1, 33, 69):
        one = psfrec.psf_muse(psd1[0], lam[[k]])[0]
        assert rel_to_peak(cube[k], one) < 1e-13
    sub = psfrec.psf_muse(psd1[0], lam[:9])
    assert rel_to_peak(cube[:9], sub) < 1e-13
    ref = orc.psf_muse(psd1[0], lam[[5, 40]])
    assert_image_close(cube[5], ref[0])
    assert_image_close(cube[40], ref[1])


def test_psf_muse_random_wavelengths(psfrec, psd1):
    """The kept-frequency / mirror tables are rebuilt per wavelength set: random
    wavelengths, including ones where output pixels fall exactly on samples, against the oracle."""
    rng = np.random.default_rng(2026)
    lam = np.concatenate([rng.uniform(486, 1100, 9), [620.8, 776.0, 970.0]])   # npix = 1000, 800, 640: integer strides
    got = psfrec.psf_muse(psd1[0], lam)
    ref = orc.psf_muse(psd1[0], lam)
    for k in range(lam.size):
        assert_image_close(got[k], ref[k])


def test_psf_muse_pruned_equals_full_grid(psfrec, psd1):
    """The pruned 80x80-sample transform against the full-grid PSF resampled on the host."""
    from scipy.interpolate import interpn
    lb = 611.0
    full = psfrec.psd_to_psf(psd1[0], None, 8, lb * 1e-9)
    npx = int(orc.npixc_of(lb)[0])
    crop = full[640 - npx // 2:640 + npx // 2, 640 - npx // 2:640 + npx // 2].copy()
    crop /= crop.sum()
    np.maximum(crop, 0, out=crop)
    pos = np.mgrid[:40, :40] * npx / 40
    ref = interpn((np.arange(npx), np.arange(npx)), crop, pos.T).T
    ref /= ref.sum()
    got = psfrec.psf_muse(psd1[0], np.array([lb]))[0]
    assert_image_close(got, ref, rtol=1e-11)


def test_underflow_cut_is_invisible(psfrec):
    """Bad seeing: most of exp(-Dphi/2) is below the underflow cut (default e^-45) and the row kernel
    flushes those entries (and whole row pairs) to zero.  The result must equal the un-cut evaluation
    to FP64 rounding and the oracle to the PSF tolerance."""
    from muse_psfr_b200 import _lib
    psd = orc.simul_psd_wfm([0.36, 0.64], (100, 10000), 1.68, 19.9)
    lam = np.array([490., 640., 930.])
    ref = orc.psf_muse(psd, lam)
    ctx = psfrec.get_context()
    default_cut = ctx.info()['exp_cut']
    try:
        ctx.set_option(_lib.OPT_EXP_CUT, 1000.0)        # cut disabled
        full = psfrec.psf_muse(psd, lam)
    finally:
        ctx.set_option(_lib.OPT_EXP_CUT, default_cut)
    assert ctx.info()['exp_cut'] == default_cut == 45.0
    cut = psfrec.psf_muse(psd, lam)
    # structure-function rows far from the centre exceed the cut at 490 nm: the skip really ran
    d = ctx.get_structure_function(0)
    assert 0.5 * (2 * np.pi / 490.) ** 2 * d[:560].min() > 64
    assert rel_to_peak(cut, full) < 1e-13
    for k in range(lam.size):
        assert_image_close(cut[k], ref[k])
        assert_image_close(full[k], ref[k])


@pytest.mark.parametrize('seeing,L0,GL', [(1.68, 19.9, 0.36), (1.57, 28.0, 0.88), (0.91, 16.9, 0.71)])
def test_graded_precision_is_invisible(psfrec, seeing, L0, GL):
    """Row pairs whose every OTF entry is below e^-25 of the peak are evaluated and transformed
    in single precision, blocks below e^-20 use the single-precision exp (psfr.h,
    PSFR_OPT_F32_ROWS / PSFR_OPT_EXP_GRADE).  The result must equal the all-FP64 evaluation
    to FP64 rounding of the peak and the oracle to the PSF tolerance."""
    from muse_psfr_b200 import _lib
    psd = orc.simul_psd_wfm([GL, 1 - GL], (100, 10000), seeing, L0)
    lam = np.array([490., 640., 780., 930.])
    ref = orc.psf_muse(psd, lam)
    ctx = psfrec.get_context()
    try:
        ctx.set_option(_lib.OPT_EXP_GRADE, 1e30)
        ctx.set_option(_lib.OPT_F32_ROWS, 1e30)
        full = psfrec.psf_muse(psd, lam)
    finally:
        ctx.set_option(_lib.OPT_EXP_GRADE, 20.0)
        ctx.set_option(_lib.OPT_F32_ROWS, 25.0)
    graded = psfrec.psf_muse(psd, lam)
    # the graded paths really ran: some row has all its entries between e^-64 and e^-25 at 490 nm
    d = ctx.get_structure_function(0)
    x = 0.5 * (2 * np.pi / 490.) ** 2 * d[:641].min(axis=1)
    assert ((x >= 25) & (x < 64)).any()
    assert rel_to_peak(graded, full) < 1e-13
    for k in range(lam.size):
        assert_image_close(graded[k], ref[k])
        assert_image_close(full[k], ref[k])


def test_row_kernels_agree(psfrec, psd1):
    """The two row kernels - PSFR_OPT_ROW_KERNEL = 2 (default, csrc/psfr_hot2.cu: one 128-thread group per
    row transform, data in shared memory, pruned third pass) and 1 (csrc/psfr_hot.cu: one warp per
    transform, data in registers) - must give the same planes."""
    from muse_psfr_b200 import _lib
    ctx = psfrec.get_context()
    bad = orc.simul_psd_wfm([0.36, 0.64], (100, 10000), 1.68, 19.9)
    for psd, lam in ((psd1[0], np.array([500., 700., 900.])), (psd1[0], LBDA35), (bad[0], LBDA35[::4])):
        ref = psfrec.psf_muse(psd, lam)
        try:
            ctx.set_option(_lib.OPT_ROW_KERNEL, 1)
            got = psfrec.psf_muse(psd, lam)
        finally:
            ctx.set_option(_lib.OPT_ROW_KERNEL, 2)
        assert np.isfinite(got).all()
        assert rel_to_peak(got, ref) < 1e-13
        want = orc.psf_muse(psd, lam[:2])
        for k in range(2):
            assert_image_close(got[k], want[k])
            assert_image_close(ref[k], want[k])

def test_config5_row_kernels_agree(psfrec, psd5, golden):
    """dim 2560: the group row kernel (two interleaved 1280-point sub-transforms combined in pass 3) against
    the warp row kernel and the reference fixture."""
    from muse_psfr_b200 import _lib
    g = golden('ref_config5')
    lam = np.concatenate([g['lbda'], np.linspace(490, 930, 100)[[0, 37, 99]]])
    ctx = psfrec.get_context(dim=2560)
    got2 = psfrec.psf_muse(psd5[0], lam)
    try:
        ctx.set_option(_lib.OPT_ROW_KERNEL, 1)
        got1 = psfrec.psf_muse(psd5[0], lam)
    finally:
        ctx.set_option(_lib.OPT_ROW_KERNEL, 2)
    assert np.isfinite(got2).all()
    assert rel_to_peak(got2, got1) < 1e-13
    for k in range(len(g['lbda'])):
        assert_image_close(got2[k], g['psf_muse'][k])


def test_wavelength_below_grid_limit(psfrec, psd1):
    with pytest.raises(ValueError):
        psfrec.psf_muse(psd1[0], np.array([400.]))      # needs a 1552-pixel crop: reference fails too


# ---------------------------------------------------------------- convolve + fit (a10, a11)
def test_convolve_final_psf(psfrec, golden):
    g, go = golden('ref_config1'), golden('oracle_config1')
    got = psfrec.convolve_final_psf(LBDA35, 1.0, 0.7, 25., g['psf_muse'])
    for i in range(35):
        assert_image_close(got[i], go['conv'][i])
    ref = orc.convolve_final_psf(LBDA35[[3]], 0.6, 0.45, 11., g['psf_muse'][[3]])
    assert_image_close(psfrec.convolve_final_psf(LBDA35[[3]], 0.6, 0.45, 11., g['psf_muse'][[3]])[0], ref[0])


def test_fit_psf_cube(psfrec, golden):
    go = golden('oracle_config1')
    tab = psfrec.fit_psf_cube(LBDA35, go['conv'])
    assert_allclose(tab['fwhm'][:, 0], go['fwhm'], rtol=FIT_RTOL)
    assert_allclose(tab['n'], go['n'], rtol=FIT_RTOL)
    assert_allclose(tab['center'], go['center'], atol=1e-5)
    assert_allclose(tab['peak'], go['peak'], rtol=FIT_RTOL)
    assert_allclose(tab['lbda'], LBDA35)


def test_fit_table_every_column(psfrec, golden):
    """flux, peak and every err_* column against the oracle's restatement of mpdaf's
    Image.moffat_fit on the same images (errors: rtol 1e-4 - the oracle's covariance comes from
    MINPACK's finite-difference Jacobian, the kernel's from the analytic one)."""
    go = golden('oracle_config1')
    sel = [0, 9, 17, 26, 34]
    tab = psfrec.fit_psf_cube(LBDA35[sel], go['conv'][sel])
    ref = orc.fit_psf_cube(LBDA35[sel], go['conv'][sel])
    assert tab.colnames == ['lbda', 'center', 'flux', 'fwhm', 'n', 'peak', 'err_center', 'err_flux', 'err_fwhm',
                            'err_n', 'err_peak', 'converged']
    assert_allclose(tab['flux'], ref['flux'], rtol=FIT_RTOL)
    assert_allclose(tab['peak'], ref['peak'], rtol=FIT_RTOL)
    assert_allclose(tab['err_center'], ref['err_center'], rtol=1e-4)
    assert_allclose(tab['err_n'], ref['err_n'], rtol=1e-4)
    assert_allclose(tab['err_peak'], ref['err_peak'], rtol=1e-4)
    assert_allclose(tab['err_fwhm'][:, 0], ref['err_fwhm'], rtol=1e-4)
    assert np.array_equal(tab['err_fwhm'][:, 0], tab['err_fwhm'][:, 1])
    assert np.array_equal(tab['err_flux'], ref['err_flux'])          # mpdaf: err_e = 0 for a circular fit
    assert tab['converged'].all()
    assert (ref['err_n'] > 0).all() and (ref['err_fwhm'] > 0).all()


def test_non_converged_fit_is_reported(psfrec, caplog):
    """A fit that runs out of iterations is flagged in the table and logged, not returned silently."""
    import logging
    img = np.zeros((1, 40, 40))          # no minimum: every parameter is degenerate
    with caplog.at_level(logging.WARNING, logger='muse_psfr.psfrec'):
        tab = psfrec.fit_psf_cube([500.], img)
    assert tab['converged'][0] == 0
    assert any('did not converge' in r.message for r in caplog.records)


def test_fit_exact_moffat_known_answer(psfrec):
    """Idempotence: images that ARE Moffat profiles return their own parameters."""
    p, q = np.mgrid[:40, :40].astype(float)
    truth = [(0.07, 20.3, 19.6, 3.1, 2.4), (0.15, 19.5, 20.5, 2.2, 1.8), (0.02, 21.0, 18.2, 5.5, 3.5)]
    imgs = np.stack([orc.moffat_model(np.array(v), p, q) for v in truth])
    tab = psfrec.fit_psf_cube(np.arange(3.), imgs)
    for k, v in enumerate(truth):
        assert_allclose(tab['peak'][k], v[0], rtol=1e-9)
        assert_allclose(tab['center'][k], v[1:3], rtol=1e-9)
        assert_allclose(tab['n'][k], v[4], rtol=1e-8)
        assert_allclose(tab['fwhm'][k, 0], 0.2 * v[3] * 2 * np.sqrt(2 ** (1 / v[4]) - 1), rtol=1e-8)


def test_fit_non_square_image(psfrec):
    p, q = np.mgrid[:32, :48].astype(float)
    img = orc.moffat_model(np.array([1.0, 15.2, 25.1, 3.0, 2.5]), p, q)
    tab = psfrec.fit_psf_cube([0.], img[None])
    assert_allclose(tab['center'][0], [15.2, 25.1], rtol=1e-8)
    assert_allclose(tab['n'][0], 2.5, rtol=1e-7)


# ---------------------------------------------------------------- compute_psf end to end (a13)
def test_compute_psf_config1(psfrec, golden):
    go = golden('oracle_config1')
    tab, cube = psfrec.compute_psf(LBDA35, 1.0, 0.7, 25., verbose=False)
    for i in range(35):
        assert_image_close(cube[i], go['conv'][i])
    assert_allclose(tab['fwhm'][:, 0], go['fwhm'], rtol=FIT_RTOL)
    assert_allclose(tab['n'], go['n'], rtol=FIT_RTOL)
    assert_allclose(tab['center'], 20, atol=1e-3)
    assert tab.meta == {'SEEING': 1.0, 'GL': 0.7, 'L0': 25.}
    assert_allclose(tab['L0'], 25.)


def test_compute_psf_reference_known_answers(psfrec):
    """test_psfrec.py:121-128 / 162-170."""
    tab, _ = psfrec.reconstruct_psf(np.array([500., 700., 900.]), 1.0, 0.7, 25., verbose=False)
    assert ['%.2f' % v for v in tab['fwhm'][:, 0]] == ['0.85', '0.73', '0.62']
    assert ['%.2f' % v for v in tab['n']] == ['2.73', '2.55', '2.23']


def test_compute_psf_config3(psfrec, golden):
    """npsflin=3 with three lasers (test_psfrec.py:77-90: fwhm 0.86 at 502.9 nm)."""
    o3 = golden('oracle_config3')
    tab, cube = psfrec.compute_psf(o3['lbda'], 1.0, 0.7, 25., npsflin=3, three_lgs_mode=True, verbose=False)
    for i in range(len(o3['lbda'])):
        assert_image_close(cube[i], o3['conv'][i])
    assert_allclose(tab['fwhm'][:, 0], o3['fwhm'], rtol=FIT_RTOL)
    assert_allclose(tab['n'], o3['n'], rtol=FIT_RTOL)


def test_batch_config4_sample(psfrec, golden):
    g4, o4 = golden('ref_config4_sample'), golden('oracle_config4_sample')
    pick = g4['pick']
    from muse_psfr_b200 import _lib
    fit, cube = psfrec.compute_psf_batch(g4['lbda'], g4['seeing'][pick], g4['GL'][pick], g4['L0'][pick],
                                         h=np.stack([g4['h0'][pick], g4['h1'][pick]], axis=1))
    for k in range(len(pick)):
        for i in range(len(g4['lbda'])):
            assert_image_close(cube[k, i], o4['conv'][k, i])
    assert_allclose(fit[:, :, _lib.FIT_FWHM] * 0.2, o4['fwhm'], rtol=FIT_RTOL)
    assert_allclose(fit[:, :, _lib.FIT_N], o4['n'], rtol=FIT_RTOL)
    assert np.all(fit[:, :, _lib.FIT_ITER] > 0)


def test_parity_sweep_slice(psfrec):
    """A 16-draw slice of tools/parity_sweep.py (VERDICT r1 task 1): 12 draws of the config-4 stream + 4
    of the sweep's corners x all 35 wavelengths, default (cut + graded) and all-FP64 options, against the
    oracle run on all host cores.  The full sweep's figures are committed under profiles/."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    import parity_sweep as ps
    seeing, GL, L0, h = ps.config4_draws(12, corners=True)
    keep = np.r_[0:12, [12, 15, 16, 19]]          # (0.4, 9, 0.3), (0.4, 29, 0.95), (2.0, 9, 0.3), (2.0, 29, 0.95)
    seeing, GL, L0, h = seeing[keep], GL[keep], L0[keep], h[keep]
    ref = ps.oracle_many([(LBDA35, seeing[i], GL[i], L0[i], h[i], {}) for i in range(16)])
    ref_cube = np.stack([r[0] for r in ref])
    ref_fw, ref_n = np.stack([r[1] for r in ref]), np.stack([r[2] for r in ref])
    for name, opts in ps.OPTION_SETS.items():
        fit, cube = ps.gpu_batch(psfrec, opts, LBDA35, seeing, GL, L0, h=h)
        res = ps.compare(cube, fit, ref_cube, ref_fw, ref_n)
        assert res['finite'] and res['not_converged'] == 0, (name, res)
        assert res['img_rel'] < PSF_RTOL and res['img_peak'] < PSF_RTOL, (name, res)
        assert res['fwhm_rel'] < FIT_RTOL and res['beta_rel'] < FIT_RTOL, (name, res)


def test_batch_is_independent_of_chunking(psfrec):
    """Size-independent property at batch scale: a draw's result does not depend on which
    other draws share its launch (chunk of 64 planes vs chunks of 7)."""
    rng = np.random.default_rng(7)
    nd = 100
    s, g, l0 = rng.uniform(.4, 2, nd), rng.uniform(.3, .95, nd), rng.uniform(9, 29, nd)
    h = np.stack([rng.uniform(50, 500, nd), rng.uniform(5000, 15000, nd)], 1)
    lam = LBDA35[::6]
    from muse_psfr_b200 import _lib
    fit_a, cube_a = psfrec.compute_psf_batch(lam, s, g, l0, h=h, max_planes=64)
    assert np.isfinite(cube_a).all() and np.all(fit_a[:, :, _lib.FIT_ITER] > 0)
    psfrec.release_contexts()
    ctx = psfrec.get_context(max_planes=7, max_lambda=35)
    recs = psfrec.draw_records(s, g, l0, h)
    fit_b = np.empty_like(fit_a)
    cube_b = np.empty_like(cube_a)
    ctx.compute_batch(recs, psfrec.direction_perf(1), psfrec._lgs_positions(False), lam, out_cube=cube_b, out_fit=fit_b)
    assert np.array_equal(cube_a, cube_b)
    assert np.array_equal(fit_a, fit_b)
    psfrec.release_contexts()
    # spot-check three draws against the oracle
    for i in (0, 57, 99):
        ref_fit, ref_cube = orc.compute_psf(lam[:2], s[i], g[i], l0[i], h=tuple(h[i]))
        assert_image_close(cube_a[i, 0], ref_cube[0])
        assert_allclose(fit_a[i, :2, _lib.FIT_FWHM] * 0.2, ref_fit['fwhm'], rtol=FIT_RTOL)
        assert_allclose(fit_a[i, :2, _lib.FIT_N], ref_fit['n'], rtol=FIT_RTOL)


# ---------------------------------------------------------------- mean + refit, polynomials (a12, a13)
def test_fused_path_equals_staged_calls(psfrec):
    """psfr_compute_batch (quadrant PSD, never materialised) is bit-identical to the reference-
    shaped stage calls simul_psd_wfm -> psf_muse -> convolve_final_psf -> fit_psf_cube, which go
    through the full N x N PSD on the host."""
    lam = LBDA35[::11]
    for npsflin, three in ((1, False), (2, True)):
        seeing, GL, L0, h = 1.3, 0.55, 16., (120., 9000.)
        tab, cube = psfrec.compute_psf(lam, seeing, GL, L0, npsflin=npsflin, h=h, three_lgs_mode=three, verbose=False)
        psd = psfrec.simul_psd_wfm([GL, 1 - GL], h, seeing, L0, npsflin=npsflin, three_lgs_mode=three, verbose=False)
        staged = psfrec.convolve_final_psf(lam, seeing, GL, L0, psfrec.psf_muse(psd, lam))
        assert np.array_equal(staged, cube)
        fit = psfrec.fit_psf_cube(lam, staged)
        assert np.array_equal(fit['fwhm'], tab['fwhm']) and np.array_equal(fit['n'], tab['n'])


def test_single_draw_single_wavelength(psfrec):
    """Smallest possible call: one draw, one wavelength (every ring / counter path with one item)."""
    lam = np.array([653.])
    fit, cube = psfrec.compute_psf_batch(lam, [0.83], [0.61], [17.5])
    ref, ref_cube = orc.compute_psf(lam, 0.83, 0.61, 17.5)
    assert cube.shape == (1, 1, 40, 40)
    assert_image_close(cube[0, 0], ref_cube[0])
    from muse_psfr_b200 import _lib
    assert_allclose(fit[0, :, _lib.FIT_FWHM] * 0.2, ref['fwhm'], rtol=FIT_RTOL)


def test_device_buffers_equal_host_buffers(psfrec):
    """The C ABI takes host or device pointers: torch tensors as plain device buffers give the
    bit-identical result of the host-buffer call (and exercise the no-copy output path)."""
    import torch
    from muse_psfr_b200 import _lib
    lam = LBDA35[::9]
    s, g, l0 = np.array([0.7, 1.2, 1.6]), np.array([0.8, 0.6, 0.45]), np.array([12., 22., 27.])
    fit_h, cube_h = psfrec.compute_psf_batch(lam, s, g, l0)
    recs = np.stack([psfrec.draw_record([g[i], 1 - g[i]], (100, 10000), s[i], l0[i], 0.,
                                        alpha_tt=psfrec.tiptilt_alpha(s[i], g[i], l0[i])) for i in range(3)])
    d_recs = torch.from_numpy(recs).cuda()
    d_cube = torch.empty((3, lam.size, 40, 40), dtype=torch.float64, device='cuda')
    d_fit = torch.empty((3, lam.size, _lib.FIT_NPAR), dtype=torch.float64, device='cuda')
    ctx = psfrec.get_context(max_planes=16, max_lambda=35)
    ctx.compute_batch(d_recs, psfrec.direction_perf(1), psfrec._lgs_positions(False), lam, out_cube=d_cube,
                      out_fit=d_fit, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_cube.cpu().numpy(), cube_h)
    assert np.array_equal(d_fit.cpu().numpy(), fit_h)


def test_c_abi_argument_errors(psfrec):
    """Error behaviour of the boundary: negative codes + message, never a crash or a fallback."""
    from muse_psfr_b200 import _lib
    ctx = psfrec.get_context()
    lib = _lib.load()
    recs = np.zeros((1, _lib.DRAW_NPAR))
    recs[0, _lib.DRAW_NLAYERS] = _lib.MAX_LAYERS + 1      # more layers than the record holds
    dirs, pos, lam = _lib.f64(psfrec.direction_perf(1)), _lib.f64(psfrec._lgs_positions(False)), np.array([500.])
    out = np.empty((1, 1, 40, 40))
    rc = lib.psfr_compute_batch(ctx._h, 1, _lib.ptr(recs), 1, _lib.ptr(dirs), 4, _lib.ptr(pos), 1, _lib.ptr(lam),
                                _lib.ptr(out), None, None)
    assert rc == _lib.E_UNSUPPORTED and b'layers' in lib.psfr_last_error(ctx._h)
    rc = lib.psfr_compute_batch(ctx._h, 0, _lib.ptr(recs), 1, _lib.ptr(dirs), 4, _lib.ptr(pos), 1, _lib.ptr(lam),
                                _lib.ptr(out), None, None)
    assert rc in (_lib.E_ARG, _lib.E_CAPACITY)
    assert lib.psfr_compute_batch(ctx._h, 1, None, 1, _lib.ptr(dirs), 4, _lib.ptr(pos), 1, _lib.ptr(lam),
                                  _lib.ptr(out), None, None) == _lib.E_ARG
    assert lib.psfr_set_option(ctx._h, 99, 1.0) == _lib.E_ARG
    assert lib.psfr_set_option(ctx._h, _lib.OPT_EXP_CUT, -1.0) == _lib.E_ARG
    assert lib.psfr_set_option(ctx._h, _lib.OPT_EXP_GRADE, 0.0) == _lib.E_ARG
    assert lib.psfr_set_option(ctx._h, _lib.OPT_F32_ROWS, -3.0) == _lib.E_ARG
    assert lib.psfr_set_option(ctx._h, _lib.OPT_ROW_KERNEL, 3.0) == _lib.E_ARG
    assert lib.psfr_psf_cube(ctx._h, 1, 1, ctx.max_lambda + 1, _lib.ptr(np.full(ctx.max_lambda + 1, 600.)),
                             _lib.ptr(out), None) < 0
    with pytest.raises(_lib.PsfrError):
        ctx.psd_to_psf(0, -1.0, np.empty((1280, 1280)))


def test_config5_field_grid(psfrec):
    """dim 2560 with a 2 x 2 field grid in three-LGS mode: direction mean inside the column
    kernel on the two-transform (NF = 2) path, against the oracle."""
    kw = dict(npsflin=2, dim=2560, three_lgs_mode=True)
    psd = psfrec.simul_psd_wfm([0.6, 0.4], (200, 9000), 0.9, 20., verbose=False, **kw)
    ref_psd = orc.simul_psd_wfm([0.6, 0.4], (200, 9000), 0.9, 20., **kw)
    assert_allclose(psd, ref_psd, rtol=1e-10, atol=0)
    lam = np.array([560.])
    got = psfrec.psf_muse(psd, lam)
    ref = orc.psf_muse(ref_psd, lam)
    assert_image_close(got[0], ref[0])


def test_two_devices_in_one_process(psfrec):
    """One context per GPU inside one process (kernel attributes are per device): the second GPU
    gives the bit-identical result of the first."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    lam = LBDA35[::12]
    args = (lam, [0.9, 1.5], [0.7, 0.5], [20., 14.])
    fit0, cube0 = psfrec.compute_psf_batch(*args, device=0)
    fit1, cube1 = psfrec.compute_psf_batch(*args, device=1)
    assert np.array_equal(cube0, cube1) and np.array_equal(fit0, fit1)


def test_mean_refit(psfrec, golden):
    go = golden('oracle_config1')
    from muse_psfr_b200 import _lib
    cubes = np.ascontiguousarray(np.stack([go['conv'], go['conv'][::-1], 0.5 * go['conv']]))
    mean = np.empty((35, 40, 40))
    fit = np.empty((35, _lib.FIT_NPAR))
    psfrec.get_context().mean_refit(3, 35, cubes, mean, fit)
    assert_allclose(mean, np.mean(cubes, axis=0), rtol=1e-15)
    ref = orc.fit_psf_cube(LBDA35[[0, 20]], np.mean(cubes, axis=0)[[0, 20]])
    assert_allclose(fit[[0, 20], _lib.FIT_FWHM] * 0.2, ref['fwhm'], rtol=FIT_RTOL)
    assert_allclose(fit[[0, 20], _lib.FIT_N], ref['n'], rtol=FIT_RTOL)


def test_fit_psf_with_polynom(psfrec, golden):
    go = golden('oracle_config1')
    pol = psfrec.fit_psf_with_polynom(LBDA35, go['fwhm'], go['n'], output=1)
    assert_allclose(pol['fwhm_pol'], go['ref_fwhm_pol'], rtol=1e-8)
    assert_allclose(pol['beta_pol'], go['ref_beta_pol'], rtol=1e-8)
    assert_allclose(pol['fwhm_fit'], go['ref_fwhm_fit'], rtol=1e-10)
    assert_allclose(pol['beta_fit'], go['ref_beta_fit'], rtol=1e-10)
    assert pol['lbda_lim'] == (475, 935)
    mixed = psfrec.fit_psf_with_polynom(LBDA35, go['fwhm'], go['n'], deg=(3, 4))
    assert_allclose(mixed['fwhm_pol'], np.polyfit(orc.norm_lbda(LBDA35), go['fwhm'], 3), rtol=1e-9)
    assert_allclose(mixed['beta_pol'], np.polyfit(orc.norm_lbda(LBDA35), go['n'], 4), rtol=1e-9)


def test_kernels_were_launched(psfrec):
    psfrec.compute_psf(np.array([600.]), 0.9, 0.6, 20., verbose=False)
    ctx = psfrec.get_context()
    assert ctx.kernel_launches() > 0
    ms, n, psfs = ctx.last_hot_timing()
    assert n == 1 and psfs == 1 and ms > 0


def test_device_exp_accuracy(psfrec):
    """csrc/fast_exp.cuh (the exp of exp(-Dphi/2), psfrec.py:793-794) against numpy/libm."""
    rng = np.random.default_rng(11)
    x = np.concatenate([-rng.uniform(0, 700, 200000), -10.0 ** rng.uniform(-12, 2.8, 100000),
                        [0.0, 1e-12, -1e-300, -0.34657359027997264, -0.3465735902799727, -745.0, -1e4]])
    y = psfrec.get_context().debug_exp(x)
    ok = x > -690
    ref = np.exp(x)
    assert np.abs(y[ok] / ref[ok] - 1).max() < 4.5e-16          # <= 2 ulp
    assert np.all(y[~ok] < 1e-299) and np.all(y >= 0)            # far below any visible level
