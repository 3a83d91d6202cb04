"""B200-native backend for the PSF-reconstruction hot path of muse-psfr (same Python API)."""
import logging as _logging
import sys as _sys

__version__ = '1.0+b200.1'


def _setup_logging():
    """INFO to stdout as "[LEVEL] message" on the reference's logger name (muse_psfr/__init__.py:1-14,
    where mpdaf.log.setup_logging does the same)."""
    log = _logging.getLogger('muse_psfr')
    if not log.handlers:
        handler = _logging.StreamHandler(_sys.stdout)
        handler.setFormatter(_logging.Formatter('[%(levelname)s] %(message)s'))
        handler.setLevel(_logging.INFO)
        log.addHandler(handler)
        log.setLevel(_logging.INFO)


_setup_logging()

from .psfrec import *  # noqa: E402,F401,F403
