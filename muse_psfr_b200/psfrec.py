"""Host-side mirror of the reference's ``muse_psfr/psfrec.py`` for the PSF-reconstruction
hot path: same function names, argument meaning and error behaviour, computed by the
sm_100a CUDA library behind the C ABI of ``include/psfr.h``.

Only scalar bookkeeping happens here (per-draw constants such as r0, the 80x80
frequency tables whose 1-ulp rounding the reference's cut-off masks depend on, result
packaging).  Every array-sized computation runs on the GPU; there is no CPU fallback.
"""
import logging
import os
from math import gamma

import numpy as np

from . import _fits, _lib
from ._lib import PsfrError

__all__ = ['compute_psf_from_sparta', 'compute_psf', 'compute_psf_batch', 'reconstruct_psf', 'create_sparta_table',
           'fit_psf_with_polynom', 'simul_psd_wfm', 'psd_to_psf', 'psf_muse', 'convolve_final_psf', 'fit_psf_cube',
           'muse_intrinsic_psf', 'direction_perf', 'seeing2r01', 'pupil_mask', 'select_sparta_rows', 'FitTable',
           'PsfrError', 'MIN_L0', 'MAX_L0']

logger = logging.getLogger('muse_psfr.psfrec')   # the reference's logger name (tests assert on it)

MIN_L0 = 8    # psfrec.py:30-31
MAX_L0 = 30

_HERE = os.path.dirname(os.path.abspath(__file__))
_WIND_DIR = np.array([0.628163, -0.326497])      # psfrec.py:66
_DIM = 1280                                      # psfrec.py:955 (compute_psf's fixed grid)
_DIMS = (1280, 2560)                             # grids the CUDA library is built for
_CONTEXTS = {}
_DEFAULT_DEVICE = [None]


# --------------------------------------------------------------------------- contexts
def set_device(device):
    """Select the GPU used by the module-level functions (default: LOCAL_RANK or 0)."""
    _DEFAULT_DEVICE[0] = int(device)


def _device():
    if _DEFAULT_DEVICE[0] is None:
        _DEFAULT_DEVICE[0] = int(os.environ.get('LOCAL_RANK', '0'))
    return _DEFAULT_DEVICE[0]


def _check_dim(dim):
    if int(dim) not in _DIMS:
        raise NotImplementedError('this build supports dim in %s only (got %r)' % (_DIMS, dim))
    return int(dim)


def get_context(max_planes=None, max_lambda=None, device=None, dim=_DIM):
    """Context cache: one context per (device, dim).  ``max_planes`` / ``max_lambda`` are minimum
    capacities - ``None`` means "whatever exists" - and a cached context only ever grows, in place
    (``Context.ensure``): objects handed out earlier stay valid."""
    device = _device() if device is None else int(device)
    dim = _check_dim(dim)
    ctx = _CONTEXTS.get((device, dim))
    if ctx is None:
        ctx = _lib.Context(device=device, dim=dim, max_planes=max_planes or (16 if dim == _DIM else 4),
                           max_lambda=max(35, max_lambda or 0))
        ctx.set_geometry(*ao_frequency_tables())
        _CONTEXTS[(device, dim)] = ctx
    else:
        ctx.ensure(max_planes, max_lambda)
    return ctx


def release_contexts():
    for ctx in _CONTEXTS.values():
        ctx.close()
    _CONTEXTS.clear()


# --------------------------------------------------------------------------- scalar helpers
def seeing2r01(seeing, lbda, zenith):
    """seeing @ 0.5 microns, lambda in microns (psfrec.py:183-187)."""
    r00p5 = 0.976 * 0.5 / seeing / 4.85
    return r00p5 * (lbda * 2) ** (6 / 5) * np.cos(np.deg2rad(zenith)) ** (3 / 5)


def direction_perf(npts, field_size=60, plot=False, lgs=None, ngs=None, ax=None):
    """Grid of directions where the PSF is estimated (psfrec.py:154-180).  The plotting
    branch of the reference (matplotlib) is outside the hot path and not provided."""
    if plot:
        raise NotImplementedError('plotting is outside the B200 hot path')
    x, y = (np.mgrid[:npts, :npts] - npts // 2) * field_size / 2
    return np.array([x, y]).reshape(2, -1)


def pupil_mask(radius, width, oc=0, inverse=False):
    """Annular pupil (psfrec.py:190-203); host helper kept for API parity - the library
    builds its own pupil and telescope OTF on the device."""
    center = (width - 1) / 2
    x, y = np.ogrid[:int(width), :int(width)]
    rho = np.hypot(x - center, y - center) / radius
    mask = (rho < 1) & (rho >= oc)
    return (~mask if inverse else mask).astype(int)


def ao_frequency_tables(dimall=80, step=8. / 40):
    """f, f_x, f_y of the AO zone, formed exactly as psfrec.py:548-554 and 241-242."""
    fx = np.fft.fftfreq(int(dimall), step)[:, np.newaxis]
    fy = fx.T
    f = np.sqrt(fx ** 2 + fy ** 2)
    with np.errstate(all='ignore'):
        arg_f = fy / fx
    arg_f[0, 0] = 0
    arg_f = np.arctan(arg_f)
    return f, f * np.cos(arg_f), f * np.sin(arg_f)


def _lgs_positions(three_lgs_mode):
    if three_lgs_mode:
        poslgs = np.array([[1, 1], [-1, -1], [-1, 1]], dtype=float).T
    else:
        poslgs = np.array([[1, 1], [-1, -1], [-1, 1], [1, -1]], dtype=float).T
    return poslgs * 63.          # psfrec.py:83-93


_COEFF = [None]


def _coeff_hl(L0):
    """coeffHL(L0) from the reference's coeffL0 calibration table (psfrec.py:895-897)."""
    if _COEFF[0] is None:
        tab = np.loadtxt(os.path.join(_HERE, 'data', 'coeffL0.txt'), dtype=np.float32)
        _COEFF[0] = (np.arange(1, tab.size + 1, dtype=np.float32), tab)
    return np.interp(L0, *_COEFF[0])


def tiptilt_alpha(seeing, GL, L0):
    """Moffat alpha [px] of the residual tip-tilt kernel, beta = 2 (psfrec.py:879-905)."""
    beta_tt = 2
    seeingHL = seeing * (1 - GL) ** (3. / 5.)
    r0HL = 0.976 * 0.5 / seeingHL / 4.85
    coeffHL = _coeff_hl(L0)
    pixscale = 0.2
    fwhmTTopt = (np.sqrt(coeffHL * 0.97 * 6.88 * (.5 * 1.e-6 / (2. * np.pi)) ** 2 *
                         8 ** (-1 / 3.) * r0HL ** (-5 / 3.)) / (4.85 * 1.e-6) * 2.35 / pixscale)
    return fwhmTTopt / (2 * np.sqrt(2 ** (1. / beta_tt) - 1))


def _wind_vectors(h_arr, wind_dir):
    """Wind vector per layer [2, nl]: 12.5 m/s (12 for integer altitudes, as ``np.full_like(h, 12.5)``
    does in the reference, psfrec.py:60-61) along the reference's two hard-coded directions
    (psfrec.py:66) or along ``wind_dir`` [rad] - required beyond two layers, where the reference
    itself fails (ValueError from broadcasting at psfrec.py:594)."""
    nl = h_arr.shape[-1]
    if wind_dir is None:
        if nl > 2:
            raise ValueError('operands could not be broadcast together: the reference has wind directions for '
                             '2 layers only; pass wind_dir= (one angle [rad] per layer) for more')
        arg_v = _WIND_DIR[:nl]
    else:
        arg_v = np.asarray(wind_dir, dtype=float)
        if arg_v.shape != (nl,):
            raise ValueError('wind_dir must hold one angle [rad] per layer')
    vent = np.full_like(h_arr, 12.5)
    return np.stack([vent * np.cos(arg_v), vent * np.sin(arg_v)])


def draw_record(Cn2, h, seeing, L0, zenith=0., alpha_tt=1.0, wind_dir=None):
    """Per-draw constants of simul_psd_wfm / dsp4muse / psd_fit, evaluated with the
    reference's own scalar expressions (psfrec.py:57-66, 108, 569-571, 594, 622-625)."""
    Cn2 = np.array(Cn2, dtype=float)
    Cn2 = Cn2 / Cn2.sum()
    h_arr = np.array(h)
    if h_arr.ndim != 1 or h_arr.size != Cn2.size:
        raise ValueError('Cn2 and h must be 1-D sequences of the same length')
    if h_arr.size > _lib.MAX_LAYERS:
        raise ValueError('at most %d turbulence layers are supported' % _lib.MAX_LAYERS)
    wind = _wind_vectors(h_arr, wind_dir)
    r0ref = seeing2r01(seeing, 0.5, zenith)
    rec = np.zeros(_lib.DRAW_NPAR)
    rec[_lib.DRAW_R0] = r0ref
    rec[_lib.DRAW_L0] = L0
    cst = ((gamma(11 / 6) ** 2 / (2 * np.pi ** (11 / 3))) * (24 * gamma(6 / 5) / 5) ** (5 / 6))
    rec[_lib.DRAW_FITC] = cst * r0ref ** (-5 / 3)
    for l in range(h_arr.size):
        rec[_lib.layer_slot(l, _lib.LAYER_CPHI)] = 0.0229 * (Cn2[l] ** (-3 / 5) * r0ref) ** (-5 / 3)
        rec[_lib.layer_slot(l, _lib.LAYER_H)] = float(h_arr[l])
        rec[_lib.layer_slot(l, _lib.LAYER_WX)] = wind[0, l]
        rec[_lib.layer_slot(l, _lib.LAYER_WY)] = wind[1, l]
    rec[_lib.DRAW_ALPHA_TT] = alpha_tt
    rec[_lib.DRAW_NLAYERS] = h_arr.size
    return rec


def draw_records(seeing, GL, L0, h=(100, 10000), zenith=0., Cn2=None, wind_dir=None):
    """Vectorised ``draw_record``.  Default profile: compute_psf's two layers Cn2 = [GL, 1 - GL]
    (psfrec.py:953); ``Cn2`` ([nl] or [ndraw, nl]) replaces it.  ``h`` is one altitude per layer or an
    array [ndraw, nl].  Same expressions as the scalar version evaluated with numpy's array kernels,
    which may differ from libm's scalar pow() by a few ulp (measured <= 3 ulp on the pow chains)."""
    seeing, GL, L0 = (np.atleast_1d(np.asarray(v, dtype=float)) for v in (seeing, GL, L0))
    nd = seeing.size
    h_arr = np.array(h)
    cn2 = np.stack([GL, 1 - GL], axis=1) if Cn2 is None else np.broadcast_to(np.asarray(Cn2, dtype=float),
                                                                               (nd, np.shape(Cn2)[-1]))
    nl = cn2.shape[1]
    if h_arr.shape[-1] != nl or h_arr.ndim > 2:
        raise ValueError('operands could not be broadcast together: h must hold %d layer altitudes' % nl)
    if nl > _lib.MAX_LAYERS:
        raise ValueError('at most %d turbulence layers are supported' % _lib.MAX_LAYERS)
    wx, wy = _wind_vectors(h_arr, wind_dir)
    cn2 = cn2 / cn2.sum(axis=1)[:, None]
    r0ref = seeing2r01(seeing, 0.5, zenith)
    cst = ((gamma(11 / 6) ** 2 / (2 * np.pi ** (11 / 3))) * (24 * gamma(6 / 5) / 5) ** (5 / 6))
    recs = np.zeros((nd, _lib.DRAW_NPAR))
    recs[:, _lib.DRAW_R0] = r0ref
    recs[:, _lib.DRAW_L0] = L0
    recs[:, _lib.DRAW_FITC] = cst * r0ref ** (-5 / 3)
    for l in range(nl):
        recs[:, _lib.layer_slot(l, _lib.LAYER_CPHI)] = 0.0229 * (cn2[:, l] ** (-3 / 5) * r0ref) ** (-5 / 3)
        recs[:, _lib.layer_slot(l, _lib.LAYER_H)] = h_arr[..., l]
        recs[:, _lib.layer_slot(l, _lib.LAYER_WX)] = wx[..., l]
        recs[:, _lib.layer_slot(l, _lib.LAYER_WY)] = wy[..., l]
    recs[:, _lib.DRAW_ALPHA_TT] = tiptilt_alpha(seeing, GL, L0)
    recs[:, _lib.DRAW_NLAYERS] = nl
    return recs


# --------------------------------------------------------------------------- result table
class FitTable:
    """Column store with the columns of the reference's ``fit_psf_cube`` table
    (psfrec.py:866-870).  ``to_astropy()`` converts when astropy is installed."""

    def __init__(self, columns, meta=None):
        self._cols = dict(columns)
        self.meta = dict(meta or {})

    @property
    def colnames(self):
        return list(self._cols)

    def __len__(self):
        return len(next(iter(self._cols.values())))

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._cols[key]
        if isinstance(key, (int, np.integer)):
            return {k: v[key] for k, v in self._cols.items()}
        return FitTable({k: v[key] for k, v in self._cols.items()}, self.meta)

    def __setitem__(self, key, value):
        n = len(self) if self._cols else None
        arr = np.asarray(value)
        self._cols[key] = np.full(n, value) if arr.ndim == 0 and n is not None else arr

    def to_astropy(self):
        from astropy.table import Table
        return Table(self._cols, meta=self.meta)

    @staticmethod
    def vstack(tables):
        names = tables[0].colnames
        return FitTable({k: np.concatenate([np.atleast_1d(t[k]) for t in tables]) for k in names},
                        tables[0].meta)


def _table_from_fit(lbda, fit, pixscale=0.2):
    """Columns of the reference's fit table (psfrec.py:866-870: mpdaf Moffat2D attributes minus
    ima / rot / cont / err_rot / err_cont, fwhm and err_fwhm x 0.2) plus ``converged``, an extension:
    0 where the Levenberg-Marquardt iteration ran out of steps (PSFR_FIT_ITER < 0); such rows are
    also reported by a warning on the module logger."""
    fwhm = fit[:, _lib.FIT_FWHM] * pixscale
    efw = fit[:, _lib.FIT_ERR_FWHM] * pixscale
    converged = fit[:, _lib.FIT_ITER] > 0
    if not converged.all():
        bad = np.flatnonzero(~converged)
        logger.warning('Moffat fit did not converge for %d of %d images (first at index %d)',
                       bad.size, len(fit), bad[0])
    return FitTable({
        'lbda': np.asarray(lbda, dtype=float),
        'center': fit[:, [_lib.FIT_Y0, _lib.FIT_X0]].copy(),
        'flux': fit[:, _lib.FIT_FLUX].copy(),
        'fwhm': np.stack([fwhm, fwhm], axis=1),
        'n': fit[:, _lib.FIT_N].copy(),
        'peak': fit[:, _lib.FIT_PEAK].copy(),
        'err_center': fit[:, [_lib.FIT_ERR_Y0, _lib.FIT_ERR_X0]].copy(),
        'err_flux': fit[:, _lib.FIT_ERR_FLUX].copy(),
        'err_fwhm': np.stack([efw, efw], axis=1),
        'err_n': fit[:, _lib.FIT_ERR_N].copy(),
        'err_peak': fit[:, _lib.FIT_ERR_PEAK].copy(),
        'converged': converged.astype(np.int64),
    })


# --------------------------------------------------------------------------- reference API
def simul_psd_wfm(Cn2, h, seeing, L0, zenith=0., plot=False, npsflin=1, dim=1280,
                  three_lgs_mode=False, verbose=True, wind_dir=None):
    """Residual-phase PSD per field direction, [npsflin**2, dim, dim] in nm^2
    (psfrec.py:36-151), synthesised on the GPU.  Extension: up to 8 layers when ``wind_dir`` gives
    one wind direction [rad] per layer (without it more than 2 layers raise ValueError, as the
    reference does)."""
    dim = _check_dim(dim)
    if three_lgs_mode and verbose:
        logger.info('Using three lasers mode')
    rec = draw_record(Cn2, h, seeing, L0, zenith, wind_dir=wind_dir)
    dirs = direction_perf(npsflin)
    ctx = get_context(max_planes=max(16 if dim == _DIM else 4, dirs.shape[1]), dim=dim)
    out = np.empty((dirs.shape[1], dim, dim))
    ctx.psd(rec[None], dirs, _lgs_positions(three_lgs_mode), out=out)
    return out


def _standard_pupil_ok(pup, dim):
    return pup is None or (np.shape(pup) == (dim // 2, dim // 2))


def psd_to_psf(psd, pup, D, lbda, phase_static=None, samp=None, FoV=None, return_all=False):
    """PSF from a residual-phase PSD (psfrec.py:689-807).  Only the branch the reference
    itself exercises is provided: MUSE pupil of dim/2 pixels, D = 8 m,
    samp = 2, FoV equal to the numerical field, no static phase."""
    psd = np.ascontiguousarray(psd, dtype=np.float64)
    dim = psd.shape[0]
    if psd.ndim != 2 or dim not in _DIMS or psd.shape[1] != dim:
        raise NotImplementedError('psd must be a square array of size %s' % (_DIMS,))
    if phase_static is not None or return_all:
        raise NotImplementedError('static phase / return_all are not on the hot path')
    if samp is not None and samp != 2:
        raise NotImplementedError("FIXME: use gridddata or spline ?")   # the reference fails here too (:640)
    if D != 8 or not _standard_pupil_ok(pup, dim):
        raise NotImplementedError('only the MUSE pupil (D=8, dim/2 pixels, oc=0.14) is supported')
    if pup is not None and not np.array_equal(np.asarray(pup) != 0, pupil_mask(dim / 4, dim / 2, oc=0.14) != 0):
        raise NotImplementedError('only the MUSE pupil (D=8, dim/2 pixels, oc=0.14) is supported')
    FoVnum = (lbda / (2 * D)) * dim / (4.85 * 1.e-6)
    if FoV is not None and not np.allclose(FoV, FoVnum):
        raise NotImplementedError("FIXME: use gridddata or spline ?")
    ctx = get_context(dim=dim)
    ctx.load_psd(psd, 1)
    ctx.structure_function(1)
    out = np.empty((dim, dim))
    ctx.psd_to_psf(0, lbda, out)
    return out


def psf_muse(psd, lambdamuse):
    """40x40 PSFs at 0.2 arcsec/pixel for each wavelength [nm] (psfrec.py:644-686)."""
    psd = np.ascontiguousarray(psd, dtype=np.float64)
    lam = np.atleast_1d(np.asarray(lambdamuse, dtype=float))
    if psd.ndim == 2:
        psd = psd[None]
    dim = psd.shape[1] if psd.ndim == 3 else 0
    if psd.ndim != 3 or dim not in _DIMS or psd.shape[2] != dim:
        raise NotImplementedError('psd must be [ndir, dim, dim] with dim in %s' % (_DIMS,))
    ndir = psd.shape[0]
    ctx = get_context(max_planes=max(16 if dim == _DIM else 4, ndir), max_lambda=max(35, lam.size), dim=dim)
    ctx.load_psd(psd, ndir)
    ctx.structure_function(ndir)
    out = np.empty((lam.size, _lib.PSF_DIM, _lib.PSF_DIM))
    try:
        ctx.psf_cube(1, ndir, lam, out)
    except PsfrError as exc:
        if exc.code == _lib.E_UNSUPPORTED:
            raise ValueError(str(exc))      # the reference raises ValueError from interpn / crop here
        raise
    return out


def muse_intrinsic_psf(lbda):
    """MUSE PSF polynomial approximation (psfrec.py:1144-1171): fwhm, beta, fwhm_std, beta_std."""
    pol_beta = [-0.83704697, 1.1337153, 0.0609222, -1.35581762, 1.15237178, 2.2106042]
    pol_fwhm = [0.60467385, -1.58905792, 1.75293264, -1.0368302, 0.21487023, 0.34851139]
    pol_beta_std = [0.18187424, -0.17841793, 0.30962616]
    pol_fwhm_std = [0.00707504, -0.0303464, 0.04596354]
    lb = (10 * np.asarray(lbda, dtype=float) - 4750) / (9350 - 4750)
    return (np.polyval(pol_fwhm, lb), np.polyval(pol_beta, lb),
            np.polyval(pol_fwhm_std, lb), np.polyval(pol_beta_std, lb))


def convolve_final_psf(lbda, seeing, GL, L0, psf):
    """Convolve with the tip-tilt and MUSE Moffat kernels (psfrec.py:874-930)."""
    lam = np.atleast_1d(np.asarray(lbda, dtype=float))
    psf = np.ascontiguousarray(psf, dtype=np.float64)
    if psf.shape != (lam.size, _lib.PSF_DIM, _lib.PSF_DIM):
        raise NotImplementedError('psf must be [nl, 40, 40]')
    ctx = get_context(max_lambda=max(35, lam.size))
    out = np.empty_like(psf)
    ctx.convolve(1, lam, [tiptilt_alpha(seeing, GL, L0)], psf, out)
    return out


def fit_psf_cube(lbda, psfcube):
    """Fit a Moffat PSF on each wavelength plane (psfrec.py:861-871).  ``psfcube`` is an
    array [nl, ny, nx] (or anything with a ``.data`` array, like an mpdaf Cube)."""
    data = getattr(psfcube, 'data', psfcube)
    data = np.ascontiguousarray(np.ma.getdata(data), dtype=np.float64)
    lam = np.atleast_1d(np.asarray(lbda, dtype=float))
    nimg, ny, nx = data.shape
    ctx = get_context()
    fit = np.empty((nimg, _lib.FIT_NPAR))
    ctx.moffat_fit(nimg, ny, nx, data, fit)
    return _table_from_fit(lam, fit)


def compute_psf(lbda, seeing, GL, L0, npsflin=1, h=(100, 10000), three_lgs_mode=False, verbose=True,
                dim=_DIM, zenith=0., Cn2=None, wind_dir=None):
    """Reconstruct a PSF from seeing, GL and L0 (psfrec.py:933-978): returns (table, psf).
    Extensions of this backend (the reference hard-codes them, psfrec.py:953-955): ``dim`` (1280 or
    2560), ``zenith`` [deg], and a turbulence profile ``Cn2`` / ``h`` / ``wind_dir`` of up to 8 layers
    in place of [GL, 1 - GL] (GL still sets the tip-tilt kernel, psfrec.py:881)."""
    if verbose:
        logger.info('Compute PSF with seeing=%.2f GL=%.2f L0=%.2f', seeing, GL, L0)
        if three_lgs_mode:
            logger.info('Using three lasers mode')
    lam = np.atleast_1d(np.asarray(lbda, dtype=float))
    fit, cube = compute_psf_batch(lam, [seeing], [GL], [L0], npsflin=npsflin, h=h,
                                  three_lgs_mode=three_lgs_mode, dim=dim, zenith=zenith, Cn2=Cn2,
                                  wind_dir=wind_dir)
    res = _table_from_fit(lam, fit[0])
    res.meta.update({'SEEING': seeing, 'GL': GL, 'L0': L0})
    res['SEEING'] = seeing
    res['GL'] = GL
    res['L0'] = L0
    return res, cube[0]


reconstruct_psf = compute_psf   # pre-1.0 name used by BASELINE.json's north_star (CHANGELOG:6)


def compute_psf_batch(lbda, seeing, GL, L0, npsflin=1, h=(100, 10000), three_lgs_mode=False,
                      out_cube=None, out_fit=None, want_cube=True, device=None, max_planes=None, stream=None,
                      dim=_DIM, zenith=0., Cn2=None, wind_dir=None):
    """Batched ``compute_psf``: one fused device pipeline for many (seeing, GL, L0[, h]) draws.

    ``h`` is one (h0, h1) pair or an array [ndraw, 2].  Returns (fit [ndraw, nl, FIT_NPAR],
    cube [ndraw, nl, 40, 40]); ``out_cube`` / ``out_fit`` may be preallocated numpy arrays or
    torch tensors (host-pinned or device) used as plain buffers."""
    lam = np.atleast_1d(np.asarray(lbda, dtype=float))
    seeing, GL, L0 = (np.atleast_1d(np.asarray(v, dtype=float)) for v in (seeing, GL, L0))
    nd = seeing.size
    if GL.size != nd or L0.size != nd:
        raise ValueError('seeing, GL and L0 must have one entry per draw (got %d, %d, %d)' % (nd, GL.size, L0.size))
    if nd == 0 or lam.size == 0:
        # nothing to compute: empty results of the right shape, written through nothing
        return (np.empty((nd, lam.size, _lib.FIT_NPAR)),
                np.empty((nd, lam.size, _lib.PSF_DIM, _lib.PSF_DIM)) if want_cube else None)
    if nd <= 64:
        # scalar expressions, bit-identical to the reference's own scalars
        h_arr = np.array(h)
        prof = None if Cn2 is None else np.asarray(Cn2, dtype=float)
        recs = np.stack([draw_record([GL[i], 1 - GL[i]] if prof is None else (prof[i] if prof.ndim == 2 else prof),
                                     h_arr[i] if h_arr.ndim == 2 else h_arr, seeing[i], L0[i], zenith,
                                     alpha_tt=tiptilt_alpha(seeing[i], GL[i], L0[i]), wind_dir=wind_dir)
                         for i in range(nd)])
    else:
        recs = draw_records(seeing, GL, L0, h, zenith=zenith, Cn2=Cn2, wind_dir=wind_dir)
    dirs = direction_perf(npsflin)
    dim = _check_dim(dim)
    if max_planes is None:
        max_planes = 128 if dim == _DIM else 16      # planes per chunk of the fused pipeline (~8 GB of workspace)
    ctx = get_context(max_planes=max(dirs.shape[1], min(max_planes, nd * dirs.shape[1])),
                      max_lambda=max(35, lam.size), device=device, dim=dim)
    if out_fit is None:
        out_fit = np.empty((nd, lam.size, _lib.FIT_NPAR))
    if out_cube is None and want_cube:
        out_cube = np.empty((nd, lam.size, _lib.PSF_DIM, _lib.PSF_DIM))
    try:
        ctx.compute_batch(recs, dirs, _lgs_positions(three_lgs_mode), lam, out_cube=out_cube,
                          out_fit=out_fit, stream=stream)
    except PsfrError as exc:
        if exc.code == _lib.E_UNSUPPORTED:
            raise ValueError(str(exc))
        raise
    return out_fit, out_cube


def _norm_lbda(lbda, lb1, lb2):
    return (lbda - lb1) / (lb2 - lb1) - 0.5


def fit_psf_with_polynom(lbda, fwhm, beta, deg=(5, 5), output=0):
    """Fit MUSE PSF fwhm and beta with polynomials in the normalised wavelength
    (psfrec.py:1174-1210); least squares solved on the device (Householder QR)."""
    lbda = np.asarray(lbda, dtype=float)
    ctx = get_context()
    if deg[0] == deg[1]:
        coef = ctx.polyfit(lbda, deg[0], np.stack([np.asarray(fwhm, float), np.asarray(beta, float)]))
        fwhm_pol, beta_pol = coef[0], coef[1]
    else:
        fwhm_pol = ctx.polyfit(lbda, deg[0], np.asarray(fwhm, float))[0]
        beta_pol = ctx.polyfit(lbda, deg[1], np.asarray(beta, float))[0]
    res = dict(fwhm_pol=fwhm_pol, beta_pol=beta_pol, lbda=lbda, lbda_lim=(475, 935))
    if output > 0:
        lbda_fit = np.linspace(475, 935, 50)
        lbf = _norm_lbda(lbda_fit, 475, 935)
        res['lbda_fit'] = lbda_fit
        res['fwhm_fit'] = np.polyval(fwhm_pol, lbf)
        res['beta_fit'] = np.polyval(beta_pol, lbf)
    return res


# --------------------------------------------------------------------------- SPARTA shell
def select_sparta_rows(values, mean_of_lgs=True, verbose=False):
    """Row rejection and laser averaging of ``compute_psf_from_sparta`` (psfrec.py:1041-1076).

    ``values``: [nrows, 4, 3] = (SEEING, TUR_GND, L0) of the 4 lasers.  Returns a list of
    (seeing, GL, L0, three_lgs_mode, row_idx, lgs_idx) and logs the reference's messages."""
    values = np.asarray(values, dtype=float)
    nrows = len(values)
    jobs = []
    for irow, v in enumerate(values, start=1):
        # GL > 0 and MIN_L0 < L0 < MAX_L0: "apparently the 4th value is often crap" (psfrec.py:1047-1051)
        ok = (v[:, 1] > 0) & (v[:, 2] < MAX_L0) & (v[:, 2] > MIN_L0)
        nb_gs = int(ok.sum())
        three_lgs_mode = nb_gs < 4
        if nb_gs == 0:
            if verbose:
                logger.info('%d/%d : No valid values, skipping this row', irow, nrows)
                logger.debug('Values: %s', v.tolist())
            continue
        elif nb_gs < 4:
            if verbose:
                logger.info('%d/%d : Using only %d values out of 4 after outliers '
                            'rejection', irow, nrows, nb_gs)
        if mean_of_lgs:
            seeing, GL, L0 = v[ok].mean(axis=0)
            jobs.append((seeing, GL, L0, three_lgs_mode, irow, -1))
        else:
            for i in np.where(ok)[0]:
                jobs.append((v[i, 0], v[i, 1], v[i, 2], three_lgs_mode, irow, i + 1))
    return jobs


def _fit_columns(lam, fit):
    """Columns of the reference's fit table for fit records [n, nl, FIT_NPAR] (flattened)."""
    nl = lam.size
    fit = np.asarray(fit).reshape(-1, _lib.FIT_NPAR)
    tab = _table_from_fit(np.tile(lam, len(fit) // nl), fit)
    return {k: tab[k] for k in tab.colnames}


def compute_psf_from_sparta(filename, extname='SPARTA_ATM_DATA', npsflin=1, lmin=490, lmax=930, nl=35,
                            lbda=None, h=(100, 10000), n_jobs=-1, plot=False, mean_of_lgs=True,
                            verbose=True, device=None):
    """Reconstruct a PSF from SPARTA data (psfrec.py:981-1120).

    ``filename``: path of a FITS file, a binary file object, or an HDUList (this package's or
    astropy's) holding the SPARTA table.  Every valid row (or laser) becomes one draw of a single
    batched GPU call (the reference fans them out over joblib processes, psfrec.py:1082-1083;
    ``n_jobs`` is accepted and only reported).  Returns an HDUList [PRIMARY, <extname>, FIT_ROWS,
    FIT_MEAN, PSF_MEAN] - an ``astropy.io.fits.HDUList`` when astropy is installed, otherwise
    the equivalent object of ``muse_psfr_b200._fits`` - or None when no row is valid."""
    if plot:
        raise NotImplementedError('plotting is outside the B200 hot path')
    hdul = _fits.from_any(filename, only={extname.upper()})
    sparta = hdul[extname]
    tbl = sparta.data
    out = _fits.HDUList([_fits.PrimaryHDU(), sparta.copy()])
    nrows = len(tbl)
    if nrows == 1:
        n_jobs = 1
    if lbda is None:
        lbda = np.linspace(lmin, lmax, nl)
    lbda = np.atleast_1d(np.asarray(lbda, dtype=float))
    if verbose:
        logger.info('Processing SPARTA table with %d values, njobs=%d ...', nrows, n_jobs)
    values = np.array([[[row['LGS%d_%s' % (k, col)] for col in ('SEEING', 'TUR_GND', 'L0')]
                        for k in range(1, 5)] for row in tbl], dtype=float).reshape(nrows, 4, 3)
    jobs = select_sparta_rows(values, mean_of_lgs=mean_of_lgs, verbose=verbose)
    if len(jobs) == 0:
        logger.warning('No valid values')
        return None

    seeing, GL, L0, three, row_idx, lgs_idx = (np.array(c) for c in zip(*jobs))
    if verbose:
        for j in range(len(jobs)):      # the messages compute_psf / simul_psd_wfm log per draw
            logger.info('Compute PSF with seeing=%.2f GL=%.2f L0=%.2f', seeing[j], GL[j], L0[j])
            if three[j]:
                logger.info('Using three lasers mode')
    nj = len(jobs)
    fit = np.empty((nj, lbda.size, _lib.FIT_NPAR))
    cube = np.empty((nj, lbda.size, _lib.PSF_DIM, _lib.PSF_DIM))
    for mode in (False, True):          # the LGS geometry is a per-call constant: 4-LGS and 3-LGS batches
        sel = np.where(three == mode)[0]
        if sel.size:
            f, c = compute_psf_batch(lbda, seeing[sel], GL[sel], L0[sel], npsflin=npsflin, h=h,
                                     three_lgs_mode=bool(mode), device=device)
            fit[sel], cube[sel] = f, c

    # fit values of all rows in one table (psfrec.py:1086-1101)
    cols = _fit_columns(lbda, fit)
    rep = lbda.size
    cols['SEEING'] = np.repeat(seeing, rep)
    cols['GL'] = np.repeat(GL, rep)
    cols['L0'] = np.repeat(L0, rep)
    cols['row_idx'] = np.repeat(np.arange(1, nj + 1, dtype=np.int64), rep)
    cols['lgs_idx'] = np.repeat(lgs_idx.astype(np.int64), rep)
    out.append(_fits.table_to_hdu(cols, name='FIT_ROWS'))

    # mean PSF over the draws, refit, median seeing / GL / L0 (psfrec.py:1104-1113)
    ctx = get_context(max_lambda=max(35, lbda.size), device=device)
    psftot = np.empty((lbda.size, _lib.PSF_DIM, _lib.PSF_DIM))
    fit_mean = np.empty((lbda.size, _lib.FIT_NPAR))
    ctx.mean_refit(nj, lbda.size, cube, psftot, fit_mean)
    med = np.median(np.stack([seeing, GL, L0], axis=1), axis=0)
    meta = {'SEEING': float(med[0]), 'GL': float(med[1]), 'L0': float(med[2])}
    out.append(_fits.table_to_hdu(_fit_columns(lbda, fit_mean[None]), meta=meta, name='FIT_MEAN'))
    out.append(_fits.ImageHDU(data=psftot, name='PSF_MEAN'))
    try:
        return _fits.to_astropy(out)
    except ImportError:
        return out


def create_sparta_table(nlines=1, seeing=1, L0=25, GL=0.7, bad_l0=False, outfile=None):
    """Helper function to create a SPARTA table with the given seeing, L0, and GL values
    (psfrec.py:1123-1141).  Returns the table HDU; ``outfile`` may be a path or a file object."""
    cols = {}
    for k in range(1, 5):
        for col, v in (('SEEING', seeing), ('TUR_GND', GL), ('L0', L0)):
            cols['LGS%d_%s' % (k, col)] = np.full(nlines, float(v))
    if bad_l0:
        cols['LGS4_L0'] = np.full(nlines, 150.0)
    hdu = _fits.table_to_hdu(cols, name='SPARTA_ATM_DATA')
    if outfile is not None:
        hdu.writeto(outfile, overwrite=True)
    return hdu
