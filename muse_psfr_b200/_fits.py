"""Minimal FITS reader / writer for the SPARTA tables and the result file of the shell
(``compute_psf_from_sparta``, the ``muse-psfr`` CLI).

The reference does its I/O with ``astropy.io.fits`` / ``astropy.table`` (psfrec.py:21-23,
1016-1023, 1095-1113, 1132-1139; cli.py:47-55, 120-121).  astropy is not part of this image,
so the shell carries this small stand-in covering exactly what that code needs: primary
header, binary tables with scalar or fixed-shape numeric columns, image extensions, header
keywords (incl. ``HIERARCH``), reading from a path or a file object, ``writeto``.  It mirrors
the slice of the astropy API the reference and its tests touch (``HDUList`` indexing by name,
``hdu.data['col']``, ``hdu.header['KEY']``, ``hdu.name``, ``.copy()``, ``.writeto()``).
``to_astropy()`` / ``from_any()`` convert at the boundary when astropy is installed.

Host-side data-format code only; nothing here is on the GPU hot path.
"""
import gzip
import io
import os
from collections import OrderedDict

import numpy as np

BLOCK = 2880
_TFORM_DTYPES = {'L': 'i1', 'B': 'u1', 'I': '>i2', 'J': '>i4', 'K': '>i8', 'E': '>f4', 'D': '>f8', 'A': 'S'}
_DTYPE_TFORM = {'f8': 'D', 'f4': 'E', 'i8': 'K', 'i4': 'J', 'i2': 'I', 'u1': 'B', 'b1': 'L'}
_BITPIX = {8: 'u1', 16: '>i2', 32: '>i4', 64: '>i8', -32: '>f4', -64: '>f8'}


class Header(OrderedDict):
    """Ordered keyword -> value map with per-keyword comments (``astropy.io.fits.Header`` subset)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.comments = {}

    def remove(self, key, ignore_missing=False):
        if key in self:
            del self[key]
        elif not ignore_missing:
            raise KeyError(key)

    def copy(self):
        h = Header(self)
        h.comments = dict(self.comments)
        return h


def _format_value(v):
    if isinstance(v, (bool, np.bool_)):
        return '%20s' % ('T' if v else 'F')
    if isinstance(v, (int, np.integer)):
        return '%20d' % int(v)
    if isinstance(v, (float, np.floating)):
        if not np.isfinite(v):
            raise ValueError('FITS headers cannot hold %r' % v)
        s = repr(float(v)).upper()
        if '.' not in s and 'E' not in s:
            s += '.0'
        return '%20s' % s
    s = str(v).replace("'", "''")
    return "'%-8s'" % s


def _card(key, value, comment=None):
    key = key.upper() if len(key) <= 8 and ' ' not in key else key
    if len(key) <= 8 and ' ' not in key:
        text = '%-8s= %s' % (key, _format_value(value))
    else:
        text = 'HIERARCH %s = %s' % (key, _format_value(value).strip())
    if comment:
        text += ' / ' + comment
    if len(text) > 80:
        if isinstance(value, str):
            raise ValueError('header card too long: %r' % text)
        text = text[:80]
    return text.ljust(80)


def _parse_value(text):
    text = text.strip()
    if not text:
        return None
    if text[0] == "'":
        out, i = [], 1
        while i < len(text):
            if text[i] == "'":
                if i + 1 < len(text) and text[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(text[i])
            i += 1
        return ''.join(out).rstrip()
    val = text.split('/')[0].strip()
    if val == 'T':
        return True
    if val == 'F':
        return False
    try:
        return int(val)
    except ValueError:
        pass
    try:
        return float(val.replace('D', 'E'))
    except ValueError:
        return val


def _parse_header(raw):
    hdr = Header()
    for i in range(0, len(raw), 80):
        card = raw[i:i + 80].decode('ascii', 'replace')
        key = card[:8].strip()
        if key == 'END':
            break
        if key == 'HIERARCH' and '=' in card:
            k, _, rest = card[8:].partition('=')
            hdr[k.strip()] = _parse_value(rest)
        elif card[8:10] == '= ':
            hdr[key] = _parse_value(card[10:])
    return hdr


def _pad(n):
    return (-n) % BLOCK


class _HDU:
    def __init__(self, data=None, header=None, name=None):
        self.data = data
        self.header = header.copy() if header is not None else Header()
        if name is not None:
            self.name = name

    @property
    def name(self):
        return str(self.header.get('EXTNAME', '')).upper()

    @name.setter
    def name(self, value):
        self.header['EXTNAME'] = str(value).upper()

    def copy(self):
        data = None if self.data is None else self.data.copy()
        new = type(self)(data=data, header=self.header)
        # an extension read from a file travels verbatim (header cards such as TUNIT / TNULL / TSCAL and
        # column types this reader only maps approximately, e.g. logical 'L', would not survive a rebuild)
        new._raw = getattr(self, '_raw', None)
        return new

    def writeto(self, fileobj, overwrite=False):
        hdus = [self] if isinstance(self, PrimaryHDU) else [PrimaryHDU(), self]
        HDUList(hdus).writeto(fileobj, overwrite=overwrite)


class PrimaryHDU(_HDU):
    @property
    def name(self):
        return 'PRIMARY'

    @name.setter
    def name(self, value):
        pass

    def _cards(self):
        cards = [_card('SIMPLE', True, 'conforms to FITS standard')]
        if self.data is None:
            cards += [_card('BITPIX', 8), _card('NAXIS', 0)]
        else:
            cards += _image_cards(self.data)
        cards.append(_card('EXTEND', True))
        return cards

    def _payload(self):
        return b'' if self.data is None else np.ascontiguousarray(self.data, dtype='>f8').tobytes()


def _image_cards(data):
    cards = [_card('BITPIX', -64, 'array data type'), _card('NAXIS', data.ndim, 'number of array dimensions')]
    for k, n in enumerate(reversed(data.shape), start=1):
        cards.append(_card('NAXIS%d' % k, n))
    return cards


class ImageHDU(_HDU):
    def _cards(self):
        data = np.asarray(self.data)
        return ([_card('XTENSION', 'IMAGE', 'Image extension')] + _image_cards(data) +
                [_card('PCOUNT', 0, 'number of parameters'), _card('GCOUNT', 1, 'number of groups')])

    def _payload(self):
        return np.ascontiguousarray(self.data, dtype='>f8').tobytes()


class BinTableHDU(_HDU):
    """Binary table; ``data`` is a numpy structured array (native byte order)."""

    def _columns(self):
        cols = []
        for nm in self.data.dtype.names:
            dt, shape = self.data.dtype[nm].base, self.data.dtype[nm].shape
            if dt.kind == 'S':
                cols.append((nm, '%dA' % dt.itemsize, 'S%d' % dt.itemsize, ()))
                continue
            code = _DTYPE_TFORM.get(dt.kind + str(dt.itemsize))
            if code is None:
                raise TypeError('column %s: dtype %s has no FITS binary-table form' % (nm, dt))
            n = int(np.prod(shape)) if shape else 1
            cols.append((nm, ('%d%s' % (n, code)) if n != 1 else code, _TFORM_DTYPES[code], shape))
        return cols

    def _cards(self):
        cols = self._columns()
        be = np.dtype([(nm, dt, shape) for nm, _, dt, shape in cols])
        cards = [_card('XTENSION', 'BINTABLE', 'binary table extension'), _card('BITPIX', 8, 'array data type'),
                 _card('NAXIS', 2, 'number of array dimensions'), _card('NAXIS1', be.itemsize, 'length of dimension 1'),
                 _card('NAXIS2', len(self.data), 'length of dimension 2'), _card('PCOUNT', 0, 'number of group parameters'),
                 _card('GCOUNT', 1, 'number of groups'), _card('TFIELDS', len(cols), 'number of table fields')]
        for k, (nm, tform, _, shape) in enumerate(cols, start=1):
            cards.append(_card('TTYPE%d' % k, nm))
            cards.append(_card('TFORM%d' % k, tform))
            if len(shape) > 1:
                cards.append(_card('TDIM%d' % k, '(%s)' % ','.join(str(s) for s in reversed(shape))))
        return cards

    def _payload(self):
        cols = self._columns()
        be = np.dtype([(nm, dt, shape) for nm, _, dt, shape in cols])
        out = np.empty(len(self.data), dtype=be)
        for nm in self.data.dtype.names:
            out[nm] = self.data[nm]
        return out.tobytes()


_STRUCTURAL = ('SIMPLE', 'XTENSION', 'BITPIX', 'NAXIS', 'PCOUNT', 'GCOUNT', 'TFIELDS', 'EXTEND', 'END')


def _is_structural(key):
    return (key in _STRUCTURAL or key.startswith('NAXIS') or
            any(key.startswith(p) and key[len(p):].isdigit() for p in ('TTYPE', 'TFORM', 'TDIM', 'TUNIT', 'TNULL',
                                                                       'TSCAL', 'TZERO', 'TDISP')))


class HDUList(list):
    """List of HDUs, indexable by position or (case-insensitive) extension name."""

    def __init__(self, hdus=(), fileobj=None):
        super().__init__(hdus)
        self._file = fileobj

    def __getitem__(self, key):
        if isinstance(key, str):
            for hdu in self:
                if hdu.name == key.upper():
                    return hdu
            raise KeyError("Extension %r not found." % key)
        return super().__getitem__(key)

    def __contains__(self, key):
        if isinstance(key, str):
            return any(h.name == key.upper() for h in self)
        return super().__contains__(key)

    def close(self):
        if self._file is not None:
            self._file.close()
            self._file = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def tobytes(self):
        buf = io.BytesIO()
        hdus = list(self)
        if not hdus or not isinstance(hdus[0], PrimaryHDU):
            hdus = [PrimaryHDU()] + hdus
        for k, hdu in enumerate(hdus):
            raw = getattr(hdu, '_raw', None)
            if raw is not None and k > 0:
                buf.write(raw[0])
                buf.write(raw[1] + b'\0' * _pad(len(raw[1])))
                continue
            cards = hdu._cards()
            cards += [_card(k, v, hdu.header.comments.get(k)) for k, v in hdu.header.items()
                      if not _is_structural(k) and v is not None]
            cards.append('END'.ljust(80))
            raw = ''.join(cards).encode('ascii')
            buf.write(raw + b' ' * _pad(len(raw)))
            payload = hdu._payload()
            buf.write(payload + b'\0' * _pad(len(payload)))
        return buf.getvalue()

    def writeto(self, fileobj, overwrite=False):
        data = self.tobytes()
        if hasattr(fileobj, 'write'):
            fileobj.write(data)
            return
        fileobj = os.fspath(fileobj)
        if os.path.exists(fileobj) and not overwrite:
            raise OSError("File %r already exists. If you mean to replace it then use the argument "
                          "'overwrite=True'." % fileobj)
        with io.open(fileobj, 'wb') as f:
            f.write(data)


def _read_header_block(f):
    raw = b''
    while True:
        block = f.read(BLOCK)
        if len(block) == 0 and not raw:
            return None
        if len(block) < BLOCK:
            raise OSError('truncated FITS header')
        raw += block
        for i in range(0, len(block), 80):
            if block[i:i + 8] == b'END     ':
                return raw


def _data_size(hdr):
    naxis = hdr.get('NAXIS', 0)
    if naxis == 0:
        return 0
    n = 1
    for k in range(1, naxis + 1):
        n *= hdr['NAXIS%d' % k]
    return abs(hdr['BITPIX']) // 8 * hdr.get('GCOUNT', 1) * (hdr.get('PCOUNT', 0) + n)


def _table_dtype(hdr):
    fields, native = [], []
    for k in range(1, hdr['TFIELDS'] + 1):
        name = hdr.get('TTYPE%d' % k, 'col%d' % k)
        tform = str(hdr['TFORM%d' % k]).strip()
        i = 0
        while i < len(tform) and tform[i].isdigit():
            i += 1
        rep = int(tform[:i]) if i else 1
        code = tform[i:i + 1]
        if code not in _TFORM_DTYPES:
            raise NotImplementedError('binary-table column %s: TFORM %r is not supported' % (name, tform))
        if code == 'A':
            fields.append((name, 'S%d' % max(rep, 1)))
            native.append((name, 'S%d' % max(rep, 1)))
            continue
        shape = ()
        tdim = hdr.get('TDIM%d' % k)
        if tdim:
            shape = tuple(int(s) for s in reversed(str(tdim).strip('() ').split(',')))
        elif rep != 1:
            shape = (rep,)
        dt = _TFORM_DTYPES[code]
        fields.append((name, dt, shape))
        native.append((name, np.dtype(dt).newbyteorder('='), shape))
    return np.dtype(fields), np.dtype(native)


def _build_hdu(hdr, payload, first):
    user = Header((k, v) for k, v in hdr.items() if not _is_structural(k))
    if first:
        data = None
        if hdr.get('NAXIS', 0) > 0 and payload:
            shape = tuple(hdr['NAXIS%d' % k] for k in range(hdr['NAXIS'], 0, -1))
            data = np.frombuffer(payload, dtype=_BITPIX[hdr['BITPIX']], count=int(np.prod(shape))).reshape(shape)
        return PrimaryHDU(data=data, header=user)
    xt = str(hdr.get('XTENSION', '')).strip()
    if xt == 'BINTABLE':
        if hdr.get('PCOUNT', 0) != 0:
            raise NotImplementedError('variable-length array columns are not supported')
        be, native = _table_dtype(hdr)
        if be.itemsize != hdr['NAXIS1']:
            raise OSError('binary table row size %d does not match NAXIS1=%d' % (be.itemsize, hdr['NAXIS1']))
        raw = np.frombuffer(payload, dtype=be, count=hdr['NAXIS2'])
        data = np.empty(raw.shape, dtype=native)
        for nm in native.names:
            data[nm] = raw[nm]
        return BinTableHDU(data=data, header=user)
    if xt == 'IMAGE':
        shape = tuple(hdr['NAXIS%d' % k] for k in range(hdr.get('NAXIS', 0), 0, -1))
        data = None
        if shape:
            data = np.frombuffer(payload, dtype=_BITPIX[hdr['BITPIX']], count=int(np.prod(shape))).reshape(shape)
            data = data.astype(data.dtype.newbyteorder('='))
            if 'BZERO' in hdr or 'BSCALE' in hdr:
                data = data * hdr.get('BSCALE', 1) + hdr.get('BZERO', 0)
        return ImageHDU(data=data, header=user)
    hdu = _HDU(data=None, header=user)    # unknown extension type: header only
    return hdu


def open(fileobj, only=None):    # noqa: A001 - mirrors astropy.io.fits.open
    """Read a FITS file (path, ``.gz`` path, bytes or binary file object) into an HDUList.
    ``only``: optional set of extension names whose data are parsed (others: header only)."""
    close = False
    if isinstance(fileobj, (bytes, bytearray)):
        f = io.BytesIO(fileobj)
    elif hasattr(fileobj, 'read'):
        f = fileobj
    else:
        path = os.fspath(fileobj)
        f = gzip.open(path, 'rb') if path.endswith('.gz') else io.open(path, 'rb')
        close = True
    try:
        hdus = []
        while True:
            raw = _read_header_block(f)
            if raw is None:
                break
            hdr = _parse_header(raw)
            size = _data_size(hdr)
            name = str(hdr.get('EXTNAME', '')).upper()
            want = only is None or name in only or not hdus
            if want:
                payload = f.read(size)
                if len(payload) < size:
                    raise OSError('truncated FITS data unit')
                f.read(_pad(size))
            else:
                payload = b''
                skip = size + _pad(size)
                if hasattr(f, 'seek') and not isinstance(f, gzip.GzipFile):
                    f.seek(skip, 1)
                else:
                    f.read(skip)
            if want:
                hdus.append(_build_hdu(hdr, payload, first=not hdus))
                hdus[-1]._raw = (raw, payload)      # see _HDU.copy
            else:
                hdus.append(_HDU(data=None, header=Header((k, v) for k, v in hdr.items() if not _is_structural(k))))
        if not hdus:
            raise OSError('empty or corrupt FITS file')
        return HDUList(hdus)
    finally:
        if close:
            f.close()


def getheader(fileobj, ext=0):
    """Header of one HDU (``astropy.io.fits.getheader``); data units are skipped."""
    hdul = open(fileobj, only=set())
    return hdul[ext].header


def table_to_hdu(columns, meta=None, name=None):
    """Build a BinTableHDU from a mapping name -> array (first axis = rows); scalar ``meta``
    entries become header keywords (``astropy.io.fits.table_to_hdu``)."""
    arrays = OrderedDict((k, np.asarray(v)) for k, v in columns.items())
    nrows = len(next(iter(arrays.values()))) if arrays else 0
    fields = []
    for k, a in arrays.items():
        if len(a) != nrows:
            raise ValueError('column %s has %d rows, expected %d' % (k, len(a), nrows))
        dt = a.dtype
        if dt.kind == 'i' and dt.itemsize not in (2, 4, 8):
            dt = np.dtype('i8')
        if dt.kind == 'U':
            dt = np.dtype('S%d' % max(1, dt.itemsize // 4))
        fields.append((k, dt, a.shape[1:]))
    data = np.empty(nrows, dtype=fields)
    for k, a in arrays.items():
        data[k] = a
    hdr = Header()
    for k, v in (meta or {}).items():
        if isinstance(v, (str, bool, int, float, np.integer, np.floating, np.bool_)):
            hdr[k] = v
    hdu = BinTableHDU(data=data, header=hdr)
    if name:
        hdu.name = name
    return hdu


def from_any(obj, only=None):
    """HDUList of this module from a path, file object, HDUList of this module, a single HDU,
    or an ``astropy.io.fits`` HDUList / HDU (serialised through memory)."""
    if isinstance(obj, HDUList):
        return obj
    if isinstance(obj, _HDU):
        return HDUList([obj] if isinstance(obj, PrimaryHDU) else [PrimaryHDU(), obj])
    mod = type(obj).__module__ or ''
    if mod.startswith('astropy.'):
        buf = io.BytesIO()
        if hasattr(obj, '__iter__'):
            obj.writeto(buf)
        else:
            from astropy.io import fits as afits
            afits.HDUList([afits.PrimaryHDU(), obj]).writeto(buf)
        return open(buf.getvalue(), only=only)
    return open(obj, only=only)


def to_astropy(hdul):
    """Convert an HDUList of this module into an ``astropy.io.fits.HDUList`` (needs astropy)."""
    from astropy.io import fits as afits
    return afits.HDUList.fromstring(hdul.tobytes())
