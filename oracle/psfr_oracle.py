"""CPU oracle for the muse-psfr PSF-reconstruction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``muse_psfr_b200``)
may import, call or execute this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and only as the checker / CPU baseline.

It is a plain numpy/scipy restatement of the reference algorithm
(``/root/reference/muse_psfr/psfrec.py``), every function citing the lines it
follows.  Pinning status:

* ``simul_psd_wfm``, ``psd_to_psf``, ``psf_muse``, ``muse_intrinsic_psf``,
  ``fit_psf_with_polynom`` are pinned against the reference's own functions run
  in the build container (stub loader, ``oracle/make_goldens.py``) and against
  the committed fixtures in ``tests/golden/``.
* ``convolve_final_psf`` and ``moffat_fit`` restate third-party code that is
  not in the reference tree (astropy ``Moffat2DKernel``, mpdaf
  ``Image.moffat_fit``; both un-vendored and unpinned in ``setup.cfg:29-34``).
  They reproduce every 2-decimal known answer of the reference's tests
  (``test_psfrec.py:22-30,36-44,58-69,77-90,121-128``).  Beyond 1e-2 the
  reference itself holds no fixture: **parity unpinned** for FWHM/beta past two
  decimals, and for the ``flux``/``peak``/``err_*`` columns and the absolute
  scale of the convolved PSF.
"""
import math
import os

import numpy as np
from numpy.fft import fft2, fftshift, ifft2
from scipy.interpolate import interpn
from scipy.optimize import leastsq
from scipy.signal import fftconvolve

_HERE = os.path.dirname(os.path.abspath(__file__))

# System constants of the MUSE WFM GLAO mode, psfrec.py:70-104 / 539-544 / 578-585
D_PUP = 8.0
ALT_DM = 1.0
LAMBDA_REF_UM = 0.5
N_ACT = 24.0
F_SAMP = 1000.0
DELAY_MS = 2.5
SEP_LGS = 63.0
NOISE_LGS2 = 1.0
DIM_PUP = 40
WIND_SPEED = 12.5
WIND_DIR = np.array([0.628163, -0.326497])  # psfrec.py:66
ARCMIN_PER_M = 60 / 206265  # psfrec.py:279 (arcmin * m -> phase slope factor)


def seeing2r01(seeing, lbda_um, zenith_deg):
    """psfrec.py:183-187."""
    r0_500 = 0.976 * 0.5 / seeing / 4.85
    return r0_500 * (lbda_um * 2) ** (6 / 5) * np.cos(np.deg2rad(zenith_deg)) ** (3 / 5)


def direction_perf(npts, field_size=60):
    """Field grid, psfrec.py:157-158 (plot branch out of scope)."""
    gx, gy = (np.mgrid[:npts, :npts] - npts // 2) * field_size / 2
    return np.array([gx, gy]).reshape(2, -1)


def lgs_positions(three_lgs_mode):
    """psfrec.py:86-93."""
    corners = [[1, 1], [-1, -1], [-1, 1]] if three_lgs_mode else \
        [[1, 1], [-1, -1], [-1, 1], [1, -1]]
    return np.array(corners, dtype=float).T * SEP_LGS


def pupil_mask(radius, width, oc=0.0):
    """Annular pupil, psfrec.py:190-203."""
    width = int(width)
    c = (width - 1) / 2
    yy, xx = np.ogrid[:width, :width]
    rho = np.hypot(yy - c, xx - c) / radius
    return ((rho < 1) & (rho >= oc)).astype(int)


def ao_frequency_tables(dimall=2 * DIM_PUP, step=D_PUP / DIM_PUP):
    """f, arg_f, f_x, f_y on the AO-zone grid exactly as the reference forms them
    (psfrec.py:548-554 and 241-242).  The masks below depend on 1-ulp rounding of
    f*cos(arctan(fy/fx)), so these tables are geometry constants that
    the host computes once and hands to the GPU."""
    fx = np.fft.fftfreq(int(dimall), step)[:, None]
    fy = fx.T
    f = np.sqrt(fx ** 2 + fy ** 2)
    with np.errstate(all='ignore'):
        arg_f = fy / fx
    arg_f[0, 0] = 0
    arg_f = np.arctan(arg_f)
    return f, arg_f, f * np.cos(arg_f), f * np.sin(arg_f)


def _wfs_transfer(f, f_x, f_y, pitch, strict):
    """Shack-Hartmann transfer function with the reference's cutoff mask,
    psfrec.py:251-257 (strict=False, '>=') and 430-435 (strict=True, '>').
    The mask keeps the reference's operator precedence: (A & B) | C."""
    wfs = 2j * np.pi * f * np.sinc(pitch * f_x) * np.sinc(pitch * f_y)
    fc = 1 / (2 * pitch)
    if strict:
        cut = (f != 0) & (np.abs(f_x) > fc) | (np.abs(f_y) > fc)
    else:
        cut = (f != 0) & (np.abs(f_x) >= fc) | (np.abs(f_y) >= fc)
    wfs[cut] = 0
    return wfs


def glao_reconstructor(f, f_x, f_y, pitch, pos_arcmin, sigr, h_recons):
    """LSE GLAO reconstructor W[ngs, s, s], psfrec.py:218-364 with LSE=True and a
    single reconstruction layer (the only case the reference supports, :339-341)."""
    ngs = pos_arcmin.shape[1]
    wfs = _wfs_transfer(f, f_x, f_y, pitch, strict=False)
    Mr = np.empty((ngs,) + f.shape, dtype=complex)
    for j in range(ngs):
        sx = f_x * pos_arcmin[0, j] * h_recons * 60 / 206265
        sy = f_y * pos_arcmin[1, j] * h_recons * 60 / 206265
        Mr[j] = wfs * np.exp(2j * np.pi * (sx + sy))
    back = Mr.conj() * (1 / sigr)[:, None, None]          # :310-313
    MAP = (back * Mr).sum(axis=0)                         # :316-320
    inv = np.zeros_like(MAP)
    nz = MAP != 0                                         # :339
    inv[nz] = 1 / MAP[nz]                                 # 1x1 np.linalg.inv, :349
    inv[0, 0] = 0                                         # :351-352
    return inv[None] * back                               # :358-362


def glao_residual_psd(f, f_x, f_y, pitch, pos_arcmin, beta_arcmin, sigv,
                      layer_psd, h_layers, h_dm, W, td, ti, wind):
    """Residual PSD in one direction (reconstruction + servo-lag + anisoplanatism
    + propagated noise), psfrec.py:367-525 with tempo=True, fitting=True."""
    ngs = pos_arcmin.shape[1]
    nl = h_layers.size
    wfs = _wfs_transfer(f, f_x, f_y, pitch, strict=True)
    dT = ti.max() + td                                    # :449
    bx, by = beta_arcmin
    p_dm = np.exp(2j * np.pi * h_dm * 60 / 206265 * (bx * f_x + by * f_y))   # :464-465
    pw = p_dm[None] * W                                   # :469-471 / 505-507
    err_rec = np.zeros(f.shape)
    for l in range(nl):
        acc = np.zeros(f.shape, dtype=complex)
        for j in range(ngs):
            sx = f_x * pos_arcmin[0, j] * h_layers[l] * 60 / 206265
            sy = f_y * pos_arcmin[1, j] * h_layers[l] * 60 / 206265
            lag = np.sinc(wind[0, l] * ti[j] * f_x + wind[1, l] * ti[j] * f_y)
            acc += pw[j] * (lag * wfs * np.exp(2j * (sx + sy) * np.pi))      # :437-443, 474-476
        p_beta = np.exp(2j * np.pi * (
            h_layers[l] * 60 / 206265 * (bx * f_x + by * f_y) -
            (wind[0, l] * dT * f_x + wind[1, l] * dT * f_y)))                 # :454-457
        proj = p_beta - acc                                                   # :480
        err_rec += (proj * layer_psd[l] * proj.conj()).real                   # :489-491
    err_rec[0, 0] = 0
    err_noise = (pw * sigv[:, None, None] * pw.conj()).sum(axis=0).real       # :514-517
    err_noise[0, 0] = 0
    return err_rec + err_noise                                                # :523-525


def wind_speed_for(h):
    """The reference builds the wind vector with ``np.full_like(h, 12.5)``
    (psfrec.py:60-61): for the default integer altitudes ``h=(100, 10000)`` the
    array is integer-typed and the speed is silently truncated to 12 m/s; float
    altitudes give 12.5 m/s.  Observable behaviour, kept."""
    return np.full_like(np.array(h), WIND_SPEED).astype(float)


def dsp4muse(Cn2, h, L0, r0ref, pos_arcsec, dir_arcsec, h_recons=ALT_DM, vent=None, wind_dir=None):
    """AO-zone PSD cube [ndir, 80, 80] (unshifted frequency order, transposed as the
    reference does at the end), psfrec.py:531-613."""
    f, _, f_x, f_y = ao_frequency_tables()
    pos = pos_arcsec / 60
    dirs = dir_arcsec / 60
    Cn2 = np.atleast_1d(np.asarray(Cn2, dtype=float))
    h = np.atleast_1d(np.asarray(h, dtype=float))
    if wind_dir is None:
        if h.size > 2:
            raise ValueError('the reference supports at most 2 layers (psfrec.py:66,594)')
        wind_dir = WIND_DIR[:h.size]
    wind_dir = np.asarray(wind_dir, dtype=float)      # extension: one direction per layer
    layer_psd = (0.0229 * (Cn2[:, None, None] ** (-3 / 5) * r0ref) ** (-5 / 3) *
                 (f ** 2 + (1 / L0) ** 2) ** (-11 / 6))                        # :569-571
    ngs = pos.shape[1]
    pitch = D_PUP / N_ACT
    sig2 = np.repeat(NOISE_LGS2, ngs)
    ti = 1 / np.repeat(F_SAMP, ngs)
    td = DELAY_MS * 1e-3
    if vent is None:
        vent = np.full_like(h, WIND_SPEED)
    wind = np.stack([vent * np.cos(wind_dir), vent * np.sin(wind_dir)])
    W = glao_reconstructor(f, f_x, f_y, pitch, pos, sig2, h_recons)
    out = np.empty((dirs.shape[1],) + f.shape)
    for b in range(dirs.shape[1]):
        out[b] = glao_residual_psd(f, f_x, f_y, pitch, pos, dirs[:, b], sig2,
                                   layer_psd, h, 1.0, W, td, ti, wind)
    return np.moveaxis(out, -1, -2)                                            # :613


def psd_fit(dim, L, r0, L0, fc):
    """Fitting-error PSD on the half-pixel-offset dim x dim grid, psfrec.py:616-626,
    returned already centred (the reference's fftshift of an fftshift-ed grid)."""
    dim = int(dim)
    u = (np.arange(dim) - (dim - 1) / 2) / L
    f = np.sqrt(u[:, None] ** 2 + u[None, :] ** 2)
    cst = (math.gamma(11 / 6) ** 2 / (2 * np.pi ** (11 / 3))) * (24 * math.gamma(6 / 5) / 5) ** (5 / 6)
    out = np.zeros_like(f)
    keep = f >= fc
    out[keep] = cst * r0 ** (-5 / 3) * (f[keep] ** 2 + (1 / L0) ** 2) ** (-11 / 6)
    return out


def simul_psd_wfm(Cn2, h, seeing, L0, zenith=0., npsflin=1, dim=1280, three_lgs_mode=False, wind_dir=None):
    """Residual-phase PSD [ndir, dim, dim] in nm^2, psfrec.py:36-151.  ``wind_dir`` (one angle per
    layer) lifts the reference's two-layer limit; the formulas are the reference's for any layer count."""
    Cn2 = np.array(Cn2, dtype=float)
    Cn2 = Cn2 / Cn2.sum()
    vent = wind_speed_for(h)
    h = np.array(h, dtype=float)
    pos = lgs_positions(three_lgs_mode)
    dirs = direction_perf(npsflin)
    r0ref = seeing2r01(seeing, LAMBDA_REF_UM, zenith)
    fc = 1 / (2 * (D_PUP / N_ACT))
    ao = dsp4muse(Cn2, h, L0, r0ref, pos, dirs, vent=vent, wind_dir=wind_dir)
    fit = psd_fit(dim, 2 * D_PUP, r0ref, L0, fc)
    psd = np.repeat(fit[None], ao.shape[0], axis=0)
    sl = slice(dim // 2 - DIM_PUP, dim // 2 + DIM_PUP)
    psd[:, sl, sl] = np.maximum(fit[sl, sl], fftshift(ao, axes=(1, 2)))        # :148-149
    return psd * (LAMBDA_REF_UM * 1000 / (2 * np.pi)) ** 2                     # :151


def structure_function_unit(psd, L=2 * D_PUP):
    """Wavelength-free structure function: Dphi(lbda) = (2 pi / lbda_nm)^2 * this.
    psfrec.py:717-722 with convnm factored out."""
    bg = ifft2(fftshift(psd)) * (psd.size / L ** 2)
    return fftshift(2 * (bg[0, 0].real - bg.real))


def telescope_otf(pup, dim):
    """Diffraction-limited OTF on the dim grid, psfrec.py:784-790 (live branch only)."""
    tab = np.zeros((dim, dim), dtype=complex)
    n = pup.shape[0]
    tab[:n, :n] = pup
    otf = fft2(np.abs(ifft2(tab)) ** 2)
    return fftshift(np.abs(otf) / pup.sum())


def psd_to_psf(psd, pup, D, lbda):
    """PSD -> PSF at wavelength lbda [m], psfrec.py:689-807, live branch only
    (samp == sampnum == 2, FoV == FoVnum, no static phase)."""
    dim = psd.shape[0]
    L = D * (dim / pup.shape[0])
    convnm = 2 * np.pi / (lbda * 1e9)
    bg = ifft2(fftshift(psd * convnm ** 2)) * (psd.size / L ** 2)
    dphi = fftshift(2 * (bg[0, 0].real - bg.real))
    otf = np.exp(-0.5 * dphi) * telescope_otf(pup, dim)
    psf = np.real(fftshift(ifft2(fftshift(otf))))
    return psf / psf.sum()


def npixc_of(lambdamuse, dimpsf=40, pixscale=0.2):
    """Crop width per wavelength, psfrec.py:663-664."""
    lam = np.atleast_1d(np.asarray(lambdamuse, dtype=float))
    return (np.round(((dimpsf * pixscale * 2 * 8 * 4.85 * 1000) / lam) / 2) * 2).astype(int)


def psf_muse(psd, lambdamuse, dimpsf=40):
    """PSD -> 40x40 PSF cube at 0.2 arcsec/pixel, psfrec.py:644-686."""
    psd = np.asarray(psd)
    cube = psd[None] if psd.ndim == 2 else psd
    dim = cube.shape[1]
    lam = np.atleast_1d(np.asarray(lambdamuse, dtype=float))
    pup = pupil_mask(dim / 4, dim / 2, oc=0.14)
    npix = npixc_of(lam, dimpsf)
    out = np.zeros((lam.size, dimpsf, dimpsf))
    c = dim // 2
    for i, (lb, npx) in enumerate(zip(lam * 1e-9, npix)):
        half = npx // 2
        acc = np.zeros((npx, npx))
        for plane in cube:
            full = psd_to_psf(plane, pup, 8, lb)
            acc += full[c - half:c + half, c - half:c + half]
        psf = acc / cube.shape[0] if psd.ndim == 3 else acc
        psf = psf / psf.sum()
        np.maximum(psf, 0, out=psf)
        pos = np.mgrid[:dimpsf, :dimpsf] * npx / dimpsf
        grid = np.arange(npx)
        out[i] = interpn((grid, grid), psf, pos.T, method='linear').T          # :635-641
    out /= out.sum(axis=(1, 2))[:, None, None]
    return out


def muse_intrinsic_psf(lbda):
    """Polynomial model of the MUSE instrumental Moffat, psfrec.py:1144-1171."""
    pol_beta = [-0.83704697, 1.1337153, 0.0609222, -1.35581762, 1.15237178, 2.2106042]
    pol_fwhm = [0.60467385, -1.58905792, 1.75293264, -1.0368302, 0.21487023, 0.34851139]
    lb = (10 * np.asarray(lbda, dtype=float) - 4750) / (9350 - 4750)
    return np.polyval(pol_fwhm, lb), np.polyval(pol_beta, lb)


def moffat2d_kernel(gamma, alpha, size=41):
    """astropy.convolution.Moffat2DKernel(gamma, alpha, x_size=size, y_size=size):
    Moffat2D amplitude (alpha-1)/(pi gamma^2), sampled at integer offsets
    (mode='center'), then normalised to unit sum (Kernel2D default)."""
    r = np.arange(size) - size // 2
    rr2 = (r[:, None] ** 2 + r[None, :] ** 2) / gamma ** 2
    k = (alpha - 1) / (np.pi * gamma ** 2) * (1 + rr2) ** (-alpha)
    return k / k.sum()


def coeff_hl_table():
    """coeffL0 calibration (L0 = 1..200 m), from the reference's coeffL0.fits."""
    tab = np.loadtxt(os.path.join(_HERE, 'coeffL0.txt'), dtype=np.float32)
    return np.arange(1, tab.size + 1, dtype=np.float32), tab


def tiptilt_alpha(seeing, GL, L0):
    """Moffat alpha (beta=2) of the residual tip-tilt kernel, psfrec.py:879-905."""
    seeing_hl = seeing * (1 - GL) ** (3. / 5.)
    r0_hl = 0.976 * 0.5 / seeing_hl / 4.85
    l0_ind, coeff = coeff_hl_table()
    coeff_hl = np.interp(L0, l0_ind, coeff)
    fwhm_tt = (np.sqrt(coeff_hl * 0.97 * 6.88 * (.5 * 1.e-6 / (2. * np.pi)) ** 2 *
                       8 ** (-1 / 3.) * r0_hl ** (-5 / 3.)) / (4.85 * 1.e-6) * 2.35 / 0.2)
    return fwhm_tt / (2 * np.sqrt(2 ** (1. / 2) - 1))


def convolve_final_psf(lbda, seeing, GL, L0, psf):
    """Tip-tilt then MUSE-intrinsic Moffat convolutions, psfrec.py:874-930."""
    lbda = np.atleast_1d(np.asarray(lbda, dtype=float))
    size = psf.shape[1] + (psf.shape[1] % 2 == 0)
    k_tt = moffat2d_kernel(tiptilt_alpha(seeing, GL, L0), 2, size)
    step1 = fftconvolve(psf, k_tt[None], mode='same')
    fwhm, beta = muse_intrinsic_psf(lbda)
    alpha = (fwhm / 0.2) / (2 * np.sqrt(2 ** (1. / beta) - 1))
    out = np.zeros_like(step1)
    for k in range(lbda.size):
        out[k] = fftconvolve(step1[k], moffat2d_kernel(alpha[k], beta[k], size), mode='same')
    return out


def moffat_start(img):
    """Start point of mpdaf Image.moffat_fit: peak pixel, width from Image.moments()
    (first-moment spread through the peak row/column) x 2 sqrt(2 ln 2), n = 2."""
    a = np.abs(img)
    total = a.sum()
    P, Q = np.indices(img.shape)
    p = np.argmax((Q * a).sum(axis=1) / total)
    q = np.argmax((P * a).sum(axis=0) / total)
    col = img[int(p), :]
    row = img[:, int(q)]
    wq = np.sqrt(np.abs((np.arange(col.size) - p) * col).sum() / np.abs(col).sum())
    wp = np.sqrt(np.abs((np.arange(row.size) - q) * row).sum() / np.abs(row).sum())
    fwhm0 = wp * 2 * np.sqrt(2 * np.log(2))
    ic = np.unravel_index(np.argmax(img), img.shape)
    n0 = 2.0
    a0 = fwhm0 / (2 * np.sqrt(2 ** (1 / n0) - 1))
    return np.array([img[ic], ic[0], ic[1], a0, n0], dtype=float), wq


def moffat_model(v, p, q):
    return v[0] * (1 + ((p - v[1]) / v[3]) ** 2 + ((q - v[2]) / v[3]) ** 2) ** (-v[4])


def moffat_fit(img):
    """mpdaf Image.moffat_fit(unit_center=None, unit_fwhm=None, circular=True,
    fit_back=False) as called at psfrec.py:863-865: unweighted 5-parameter
    least squares over every pixel (scipy.optimize.leastsq = MINPACK lmdif),
    re-run while centre or width moved by more than 0.1 pixel.
    Returns dict(center, fwhm[px], n, peak, flux, err_*, v)."""
    img = np.asarray(img, dtype=float)
    p, q = np.indices(img.shape)
    p = p.ravel().astype(float)
    q = q.ravel().astype(float)
    data = img.ravel()

    def resid(v):
        return moffat_model(v, p, q) - data

    v0, _ = moffat_start(img)
    v, cov, info, _, _ = leastsq(resid, v0.copy(), full_output=1)
    while abs(v[1] - v0[1]) > 0.1 or abs(v[2] - v0[2]) > 0.1 or abs(v[3] - v0[3]) > 0.1:
        v0 = v
        v, cov, info, _, _ = leastsq(resid, v0.copy(), full_output=1)
    chisq = float((info['fvec'] ** 2).sum())
    dof = data.size - v.size
    err = (np.sqrt(np.abs(np.diag(cov)) * abs(chisq / dof)) if cov is not None
           else np.full(v.size, np.nan))
    a = abs(v[3])
    n = v[4]
    k = 2 * np.sqrt(2 ** (1 / n) - 1)
    fwhm = a * k
    # Derived columns as mpdaf's Image.moffat_fit forms them for fit_n=True, circular=True (mpdaf 3.x,
    # mpdaf/obj/image.py, restated from the published source - the package is neither in the reference
    # tree nor installable here, so these columns are PARITY UNPINNED): e = 1, err_e = 0,
    #   flux = I / (n - 1) * (pi a^2 e),  err_fwhm = err_a * n,  err_flux = err_I err_n err_a^2 err_e (= 0).
    err_fwhm = err[3] * n
    err_flux = err[0] * err[4] * err[3] * err[3] * 0.0
    return dict(v=v, center=np.array([v[1], v[2]]), fwhm=fwhm, n=n, peak=v[0],
                flux=v[0] / (n - 1) * (np.pi * a * a), chisq=chisq, err_a=err[3], err_flux=err_flux,
                err_center=err[1:3], err_n=err[4], err_peak=err[0], err_fwhm=err_fwhm)


def fit_psf_cube(lbda, cube, pixscale=0.2):
    """Per-plane Moffat fit, psfrec.py:861-871: fwhm in arcsec (x0.2), n, centre [px]."""
    lbda = np.atleast_1d(np.asarray(lbda, dtype=float))
    fits = [moffat_fit(im) for im in cube]
    return dict(lbda=lbda,
                center=np.array([f['center'] for f in fits]),
                fwhm=np.array([f['fwhm'] for f in fits]) * pixscale,
                n=np.array([f['n'] for f in fits]),
                peak=np.array([f['peak'] for f in fits]),
                flux=np.array([f['flux'] for f in fits]),
                err_center=np.array([f['err_center'] for f in fits]),
                err_flux=np.array([f['err_flux'] for f in fits]),
                err_fwhm=np.array([f['err_fwhm'] for f in fits]) * pixscale,
                err_n=np.array([f['err_n'] for f in fits]),
                err_peak=np.array([f['err_peak'] for f in fits]),
                chisq=np.array([f['chisq'] for f in fits]))


def compute_psf(lbda, seeing, GL, L0, npsflin=1, h=(100, 10000), three_lgs_mode=False, dim=1280,
                zenith=0., Cn2=None, wind_dir=None):
    """Per-draw driver, psfrec.py:933-978 (returns (fit dict, psf[nl,40,40])); zenith / Cn2 / wind_dir
    are the extensions of SURVEY 8(f4) (the reference fixes zenith = 0 and Cn2 = [GL, 1 - GL])."""
    lbda = np.atleast_1d(np.asarray(lbda, dtype=float))
    psd = simul_psd_wfm([GL, 1 - GL] if Cn2 is None else Cn2, h, seeing, L0, zenith=zenith, npsflin=npsflin,
                        dim=dim, three_lgs_mode=three_lgs_mode, wind_dir=wind_dir)
    psf = psf_muse(psd[0] if npsflin == 1 else psd, lbda)
    psf = convolve_final_psf(lbda, seeing, GL, L0, psf)
    res = fit_psf_cube(lbda, psf)
    res.update(SEEING=seeing, GL=GL, L0=L0)
    return res, psf


def norm_lbda(lbda, lb1=475, lb2=935):
    """psfrec.py:1213-1215."""
    return (np.asarray(lbda, dtype=float) - lb1) / (lb2 - lb1) - 0.5


def fit_psf_with_polynom(lbda, fwhm, beta, deg=(5, 5), output=0):
    """Polynomial smoothing of fwhm(lbda), beta(lbda), psfrec.py:1174-1210."""
    lb = norm_lbda(lbda)
    res = dict(fwhm_pol=np.polyfit(lb, fwhm, deg[0]), beta_pol=np.polyfit(lb, beta, deg[1]),
               lbda=lbda, lbda_lim=(475, 935))
    if output > 0:
        grid = np.linspace(475, 935, 50)
        res['lbda_fit'] = grid
        res['fwhm_fit'] = np.polyval(res['fwhm_pol'], norm_lbda(grid))
        res['beta_fit'] = np.polyval(res['beta_pol'], norm_lbda(grid))
    return res


# --------------------------------------------------------------------------
# SPARTA-row handling used by the CPU baseline and the shell tests
# --------------------------------------------------------------------------
MIN_L0 = 8
MAX_L0 = 30


def select_sparta_rows(values, mean_of_lgs=True):
    """Row rejection and laser averaging of compute_psf_from_sparta,
    psfrec.py:1041-1076.  values: [nrows, 4, 3] (seeing, GL, L0 per laser).
    Returns list of (seeing, GL, L0, three_lgs_mode, row_idx, lgs_idx)."""
    jobs = []
    for irow, v in enumerate(np.asarray(values, dtype=float), start=1):
        ok = (v[:, 1] > 0) & (v[:, 2] < MAX_L0) & (v[:, 2] > MIN_L0)
        nb = int(ok.sum())
        if nb == 0:
            continue
        three = nb < 4
        if mean_of_lgs:
            s, g, l0 = v[ok].mean(axis=0)
            jobs.append((s, g, l0, three, irow, -1))
        else:
            for i in np.where(ok)[0]:
                jobs.append((v[i, 0], v[i, 1], v[i, 2], three, irow, i + 1))
    return jobs
