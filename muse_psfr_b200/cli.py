"""``muse-psfr`` console entry of the CUDA backend.

Behaviour pinned by the reference's tests (muse_psfr/test_psfrec.py:103-170; the reference
script is muse_psfr/cli.py:13-122): the option names, the three ``SystemExit`` messages, the
order of the INFO records, the layout of the appended log file and the HDU names of ``--outfile``.
Everything else - parser construction, report assembly, colouring - is this package's own.

    python -m muse_psfr_b200 --values 1,0.7,25 --no-color
"""
import argparse
import io
import logging
import sys

from . import __version__, _fits
from .psfrec import compute_psf_from_sparta, create_sparta_table

logger = logging.getLogger('muse_psfr.cli')

RULE = '-' * 68
# wavelengths reported by the script (nm): 3 instead of 35, the reference's choice for speed
REPORT_LBDA = dict(lmin=500, lmax=900, nl=3)

# (option strings, argparse keywords)
_OPTIONS = [
    (('raw',), dict(nargs='?', help='raw observation file holding the SPARTA_ATM_DATA table')),
    (('--values',), dict(metavar='SEEING,GL,L0', help='three comma-separated numbers used instead of a raw file')),
    (('--logfile',), dict(default='muse_psfr.log', help='text file the report is appended to')),
    (('-o', '--outfile'), dict(help='FITS file receiving the per-row fits, the mean fit and the mean PSF')),
    (('--njobs',), dict(default=-1, type=int, help='kept for compatibility; every row runs in one GPU batch')),
    (('--verbose', '-v'), dict(action='store_true', help='DEBUG-level logging')),
    (('--no-color',), dict(action='store_true', help='plain report lines')),
    (('--plot',), dict(action='store_true', help='not available in this backend (matplotlib is off the hot path)')),
    (('--version',), dict(action='version', version='%(prog)s ' + __version__)),
]

# rows of the report: label, column of FIT_MEAN, scale, number format
_ROWS = [('LBDA', 'lbda', 10.0, '%.0f'), ('FWHM', 'fwhm', 1.0, '%.2f'), ('BETA', 'n', 1.0, '%.2f')]


def _parse(argv):
    parser = argparse.ArgumentParser(description='MUSE-PSFR version %s' % __version__)
    for flags, kw in _OPTIONS:
        parser.add_argument(*flags, **kw)
    return parser.parse_args(argv)


def _telemetry(opts):
    """(source for compute_psf_from_sparta, observation title or None)."""
    if opts.values:
        try:
            triple = [float(tok) for tok in opts.values.split(',')]
        except ValueError:
            triple = []
        if len(triple) != 3:
            sys.exit('--values must contain a list of 3 comma-separated values for seeing, GL, and L0')
        seeing, gl, l0 = triple
        buf = io.BytesIO()
        create_sparta_table(seeing=seeing, GL=gl, L0=l0, outfile=buf)
        buf.seek(0)
        return buf, None
    if opts.raw is None:
        sys.exit('no input file provided')
    card = _fits.getheader(opts.raw).get
    title = 'OB %s %s Airmass %.2f-%.2f' % (card('ESO OBS NAME'), card('DATE'),
                                            card('ESO TEL AIRM START', 0), card('ESO TEL AIRM END', 0))
    return opts.raw, title


def _painter(plain):
    """Function (label, three formatted numbers) -> report line; coloured when colorama exists."""
    if not plain:
        try:
            from colorama import Back, Fore, Style
        except ImportError:
            plain = True
    if plain:
        return (lambda label, cells: ' '.join([label] + cells)), ''
    inks = (Fore.BLUE, Fore.GREEN, Fore.RED)
    lead, tail = Back.BLACK + Style.BRIGHT + Fore.WHITE, Fore.RESET + Style.NORMAL + Back.RESET

    def paint(label, cells):
        return lead + label + ''.join(' ' + ink + cell for ink, cell in zip(inks, cells)) + tail
    return paint, Style.RESET_ALL


def _report(title, fit_mean, plain):
    """Lines of the report for the FIT_MEAN HDU."""
    hdr, data = fit_mean.header, fit_mean.data
    lines = [title] if title else []
    lines += [RULE, 'Sparta Seeing: %.2f arcsec GL: %.2f L0:%.2f m' % (hdr['SEEING'], hdr['GL'], hdr['L0'])]
    paint, reset = _painter(plain)
    for label, col, scale, fmt in _ROWS:
        values = data[col]
        if getattr(values, 'ndim', 1) == 2:      # fwhm is stored per axis
            values = values[:, 0]
        lines.append(paint(label, [fmt % (scale * v) for v in values]))
    lines.append(reset + RULE)
    return lines


def main(args=None):
    opts = _parse(args)
    logger.info('MUSE-PSFR version %s', __version__)
    source, title = _telemetry(opts)
    if title:
        logger.info(title)
    logger.info('Computing PSF Reconstruction from Sparta data')
    if opts.verbose:
        top = logging.getLogger('muse_psfr')
        top.setLevel(logging.DEBUG)
        for handler in top.handlers[:1]:
            handler.setLevel(logging.DEBUG)
    if opts.plot:
        sys.exit('--plot is not available in the B200 backend')

    result = compute_psf_from_sparta(source, n_jobs=opts.njobs, **REPORT_LBDA)
    if not result:
        sys.exit('No results')

    lines = _report(title, result['FIT_MEAN'], opts.no_color)
    for line in lines:
        logger.info(line)
    if opts.logfile is not None:
        with open(opts.logfile, 'a') as fd:
            fd.write('\nFile: %s\n' % opts.raw)
            fd.write('\n'.join(lines) + '\n')
        logger.info('Results saved to %s', opts.logfile)
    if opts.outfile is not None:
        result.writeto(opts.outfile, overwrite=True)
        logger.info('FITS file saved to %s', opts.outfile)


if __name__ == '__main__':
    main()
