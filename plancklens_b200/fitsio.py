"""Minimal FITS binary-table reader / writer for HEALPix maps and alm.

healpy writes its files through astropy; neither is installed here, so this module implements the part of the FITS
standard those files use -- an empty primary HDU followed by BINTABLE extensions with fixed-width numeric columns --
in numpy, with the column names, formats and header keywords of `healpy.write_alm` / `healpy.write_map`:

  alm : columns `index` (J, l*l + l + m + 1), `real` (D), `imag` (D), one row per coefficient; MAX-LPOL / MAX-MPOL.
  map : one column per map, 1024 pixels per row (TFORM '1024D') when npix is a multiple of 1024, PIXTYPE = 'HEALPIX',
        ORDERING, NSIDE, FIRSTPIX, LASTPIX, INDXSCHM = 'IMPLICIT'.

Files ending in `.gz` are gzip-compressed.  Parity note: written from the FITS standard and the healpy conventions;
no healpy-written sample is available in this environment, so compatibility with a real healpy is unpinned -- the
tests check the block structure, the header cards and the round trip.
"""
import gzip
import re

import numpy as np

BLOCK = 2880
_TFORM = {'L': 'i1', 'B': 'u1', 'I': '>i2', 'J': '>i4', 'K': '>i8', 'E': '>f4', 'D': '>f8'}
_CODE = {np.dtype('int16'): 'I', np.dtype('int32'): 'J', np.dtype('int64'): 'K',
         np.dtype('float32'): 'E', np.dtype('float64'): 'D', np.dtype('uint8'): 'B'}


def _open(fname, mode):
    if str(fname).endswith('.gz'):
        return gzip.open(fname, mode, compresslevel=1) if 'w' in mode else gzip.open(fname, mode)
    return open(fname, mode)


def _card(key, value=None, comment=''):
    """One 80-character header card (FITS standard 4.1, fixed format)."""
    if value is None:
        return ('%-8s' % key).ljust(80)
    if isinstance(value, bool):
        v = '%20s' % ('T' if value else 'F')
    elif isinstance(value, (int, np.integer)):
        v = '%20d' % value
    elif isinstance(value, (float, np.floating)):
        v = '%20s' % ('%.16G' % value)
    else:
        v = "'%-8s'" % str(value).replace("'", "''")
        v = '%-20s' % v
    card = '%-8s= %s' % (key, v)
    if comment:
        card += ' / ' + comment
    return card[:80].ljust(80)


def _pad(b, fill):
    r = (-len(b)) % BLOCK
    return b + fill * r


def _header_bytes(cards):
    return _pad(''.join(cards + [_card('END')]).encode('ascii'), b' ')


def _primary():
    return _header_bytes([_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', 8, 'array data type'),
                          _card('NAXIS', 0, 'number of array dimensions'), _card('EXTEND', True)])


def _bintable(columns, extra=()):
    """columns: list of (name, 1-D or 2-D array, unit).  2-D arrays become vector columns (repeat = shape[1])."""
    nrow = len(columns[0][1])
    fields, cards_cols = [], []
    for i, (name, arr, unit) in enumerate(columns, 1):
        arr = np.asarray(arr)
        assert len(arr) == nrow
        rep = 1 if arr.ndim == 1 else arr.shape[1]
        code = _CODE[arr.dtype]
        fields.append(('f%d' % i, _TFORM[code], (rep,)) if rep > 1 else ('f%d' % i, _TFORM[code]))
        cards_cols += [_card('TTYPE%d' % i, name), _card('TFORM%d' % i, '%d%s' % (rep, code) if rep > 1 else code)]
        if unit:
            cards_cols.append(_card('TUNIT%d' % i, unit))
    dt = np.dtype(fields)
    rec = np.empty(nrow, dtype=dt)
    for i, (name, arr, unit) in enumerate(columns, 1):
        rec['f%d' % i] = arr
    cards = [_card('XTENSION', 'BINTABLE', 'binary table extension'), _card('BITPIX', 8, 'array data type'),
             _card('NAXIS', 2, 'number of array dimensions'), _card('NAXIS1', dt.itemsize, 'length of dimension 1'),
             _card('NAXIS2', nrow, 'length of dimension 2'), _card('PCOUNT', 0, 'number of group parameters'),
             _card('GCOUNT', 1, 'number of groups'), _card('TFIELDS', len(columns), 'number of table fields')]
    cards += cards_cols + [_card(*e) for e in extra]
    return _header_bytes(cards) + _pad(rec.tobytes(), b'\0')


def _parse_value(s):
    s = s.split('/')[0].strip() if not s.strip().startswith("'") else s.strip()
    if s.startswith("'"):
        m = re.match(r"'((?:[^']|'')*)'", s)
        return m.group(1).replace("''", "'").rstrip() if m else s
    if s in ('T', 'F'):
        return s == 'T'
    try:
        return int(s)
    except ValueError:
        try:
            return float(s.replace('D', 'E'))
        except ValueError:
            return s


def _read_header(f):
    hdr = {}
    while True:
        blk = f.read(BLOCK)
        if len(blk) < BLOCK:
            return None
        for i in range(0, BLOCK, 80):
            card = blk[i:i + 80].decode('ascii', errors='replace')
            key = card[:8].strip()
            if key == 'END':
                return hdr
            if card[8:10] == '= ':
                hdr[key] = _parse_value(card[10:])


def read_hdu(fname, hdu=1):
    """-> (header dict, structured array of the rows) of BINTABLE extension number `hdu` (1 = first extension)."""
    with _open(fname, 'rb') as f:
        k = 0
        while True:
            hdr = _read_header(f)
            if hdr is None:
                raise IOError('%s: HDU %d not found' % (fname, hdu))
            naxis = hdr.get('NAXIS', 0)
            nbytes = 0
            if naxis > 0:
                nbytes = abs(hdr['BITPIX']) // 8 * hdr.get('GCOUNT', 1)
                n = 1
                for a in range(1, naxis + 1):
                    n *= hdr['NAXIS%d' % a]
                nbytes *= (hdr.get('PCOUNT', 0) + n)
            padded = nbytes + (-nbytes) % BLOCK
            if k == hdu:
                assert hdr.get('XTENSION', '').strip() == 'BINTABLE', hdr.get('XTENSION')
                fields = []
                for i in range(1, hdr['TFIELDS'] + 1):
                    m = re.match(r'(\d*)([A-Z])', hdr['TFORM%d' % i].strip())
                    rep = int(m.group(1)) if m.group(1) else 1
                    base = _TFORM[m.group(2)]
                    fields.append(('f%d' % i, base, (rep,)) if rep > 1 else ('f%d' % i, base))
                dt = np.dtype(fields)
                assert dt.itemsize == hdr['NAXIS1'], (dt.itemsize, hdr['NAXIS1'])
                data = np.frombuffer(f.read(nbytes), dtype=dt, count=hdr['NAXIS2'])
                return hdr, data
            f.seek(padded, 1) if not isinstance(f, gzip.GzipFile) else f.read(padded)
            k += 1


# ---------------------------------------------------------------------------------------------- alm
import functools


@functools.lru_cache(maxsize=8)
def _alm_lm(L):
    """(l, m, healpy explicit index l*l + l + m + 1) of the m-major triangular layout with lmax = mmax = L"""
    m = np.repeat(np.arange(L + 1), np.arange(L + 1, 0, -1))
    start = np.concatenate([[0], np.cumsum(np.arange(L + 1, 0, -1))[:-1]])
    l = np.arange(m.size) - np.repeat(start, np.arange(L + 1, 0, -1)) + m
    index = (l * l + l + m + 1).astype(np.int32)
    for x in (l, m, index):
        x.setflags(write=False)
    return l, m, index


def write_alm(fname, alm, lmax=None, mmax=None, out_dtype=np.float64):
    """healpy.write_alm layout for one alm (m-major triangular input, mmax = lmax)."""
    alm = np.asarray(alm)
    L = int(np.floor(np.sqrt(2 * alm.size) - 1))
    assert (L + 1) * (L + 2) // 2 == alm.size, 'alm size does not match a triangular layout'
    lmax = L if lmax is None or lmax < 0 else min(lmax, L)
    mmax = lmax if mmax is None or mmax < 0 else min(mmax, lmax)
    l, m, index = _alm_lm(L)
    a = alm
    if lmax < L or mmax < lmax:
        keep = (l <= lmax) & (m <= mmax)
        index, a = index[keep], alm[keep]
    cols = [('index', index, 'l*l+l+m+1'), ('real', a.real.astype(out_dtype, copy=False), 'unknown'),
            ('imag', a.imag.astype(out_dtype, copy=False), 'unknown')]
    extra = [('MAX-LPOL', int(lmax), 'Maximum L multipole'), ('MAX-MPOL', int(mmax), 'Maximum M multipole'),
             ('EXTNAME', 'xtension', 'name of this binary table extension')]
    with _open(fname, 'wb') as f:
        f.write(_primary())
        f.write(_bintable(cols, extra))


def read_alm(fname, hdu=1, return_mmax=False):
    hdr, d = read_hdu(fname, hdu)
    idx = d['f1']
    n = idx.size
    L = int(np.floor(np.sqrt(2 * n) - 1))
    if (L + 1) * (L + 2) // 2 == n and np.array_equal(idx, _alm_lm(L)[2]):
        # the complete m-major triangle, as written by write_alm above: rows are already in array order
        alm = np.empty(n, dtype=np.complex128)
        alm.real = d['f2']
        alm.imag = d['f3']
        return (alm, L) if return_mmax else alm
    idx = idx.astype(np.int64)
    l = np.floor(np.sqrt(idx - 1)).astype(np.int64)
    m = idx - l * l - l - 1
    if np.any(m < 0) or np.any(m > l):          # guard against sqrt rounding at perfect squares
        l = np.where(m < 0, l - 1, l)
        m = idx - l * l - l - 1
    lmax, mmax = int(l.max()), int(m.max())
    size = mmax * (2 * lmax + 1 - mmax) // 2 + lmax + 1
    alm = np.zeros(size, dtype=np.complex128)
    alm[m * (2 * lmax + 1 - m) // 2 + l] = d['f2'].astype(np.float64) + 1j * d['f3'].astype(np.float64)
    return (alm, mmax) if return_mmax else alm


# ---------------------------------------------------------------------------------------------- maps
_MAP_NAMES = ['TEMPERATURE', 'Q_POLARISATION', 'U_POLARISATION']


def write_map(fname, maps, nest=False, dtype=np.float64, coord=None, column_names=None):
    """healpy.write_map layout: one BINTABLE, one column per map, 1024 pixels per row when npix allows it."""
    maps = np.asarray(maps)
    if maps.ndim == 1:
        maps = maps[None, :]
    npix = maps.shape[1]
    nside = int(round(np.sqrt(npix / 12)))
    assert 12 * nside * nside == npix, 'not a HEALPix map size'
    rep = 1024 if npix % 1024 == 0 else 1
    names = column_names or (_MAP_NAMES if len(maps) <= 3 else ['column_%d' % i for i in range(len(maps))])
    cols = [(names[i], np.ascontiguousarray(mp, dtype=dtype).reshape(-1, rep) if rep > 1 else np.ascontiguousarray(mp, dtype=dtype), '')
            for i, mp in enumerate(maps)]
    extra = [('PIXTYPE', 'HEALPIX', 'HEALPIX pixelisation'), ('ORDERING', 'NESTED' if nest else 'RING', 'Pixel ordering scheme'),
             ('EXTNAME', 'xtension', 'name of this binary table extension'), ('NSIDE', nside, 'Resolution parameter of HEALPIX'),
             ('FIRSTPIX', 0, 'First pixel # (0 based)'), ('LASTPIX', npix - 1, 'Last pixel # (0 based)'),
             ('INDXSCHM', 'IMPLICIT', 'Indexing: IMPLICIT or EXPLICIT'), ('OBJECT', 'FULLSKY', 'Sky coverage')]
    if coord:
        extra.append(('COORDSYS', coord, 'Ecliptic, Galactic or Celestial (equatorial)'))
    with _open(fname, 'wb') as f:
        f.write(_primary() + _bintable(cols, extra))


def read_map(fname, field=0, hdu=1, return_header=False):
    """-> float64 map (or list of maps for a tuple `field`) in the file's ordering, and optionally the header."""
    hdr, d = read_hdu(fname, hdu)
    one = np.isscalar(field)
    out = [np.ascontiguousarray(d['f%d' % (i + 1)]).reshape(-1).astype(np.float64) for i in ([field] if one else field)]
    ret = out[0] if one else out
    return (ret, hdr) if return_header else ret
