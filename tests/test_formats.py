"""On-disk formats: FITS binary tables for alm / maps in healpy's layout, the sqlite
`npdb` / `fldb` caches of the reference.  CPU only."""
import gzip
import sqlite3

import numpy as np
import pytest

from helpers import rand_alm


def test_fits_alm_layout_and_round_trip(tmp_path):
    from plancklens_b200 import fitsio, hp
    lmax = 37
    alm = rand_alm(np.random.default_rng(0), lmax)
    fn = str(tmp_path / 'sim_0000_tlm.fits')
    hp.write_alm(fn, alm)
    raw = open(fn, 'rb').read()
    assert len(raw) % 2880 == 0 and raw[:30] == b'SIMPLE  =                    T'
    hdr, d = fitsio.read_hdu(fn, 1)
    assert hdr['XTENSION'] == 'BINTABLE' and hdr['TFIELDS'] == 3 and hdr['NAXIS1'] == 20 and hdr['NAXIS2'] == alm.size
    assert (hdr['TTYPE1'], hdr['TFORM1'], hdr['TUNIT1']) == ('index', 'J', 'l*l+l+m+1')
    assert (hdr['TTYPE2'], hdr['TFORM2'], hdr['TTYPE3'], hdr['TFORM3']) == ('real', 'D', 'imag', 'D')
    assert hdr['MAX-LPOL'] == lmax and hdr['MAX-MPOL'] == lmax
    # healpy's explicit index: l^2 + l + m + 1, rows in the m-major order of the array
    l, m = hp.Alm.getlm(lmax)
    assert np.array_equal(d['f1'], l * l + l + m + 1)
    # big-endian doubles on disk: the first data row starts right after two header blocks
    first = np.frombuffer(raw[2 * 2880:2 * 2880 + 20], dtype=np.dtype([('i', '>i4'), ('r', '>f8'), ('c', '>f8')]))
    assert first['i'][0] == 1 and first['r'][0] == alm[0].real
    back, mmax = hp.read_alm(fn, return_mmax=True)
    assert mmax == lmax and np.array_equal(back, alm)
    # truncation on write, as healpy's lmax / mmax arguments
    hp.write_alm(fn, alm, lmax=20)
    b20 = hp.read_alm(fn)
    assert b20.size == hp.Alm.getsize(20)
    assert np.array_equal(b20, np.concatenate([alm[hp.Alm.getidx(lmax, np.arange(mm, 21), mm)] for mm in range(21)]))
    with pytest.raises(OSError):
        hp.write_alm(fn, alm, overwrite=False)


@pytest.mark.parametrize("nside,gz", [(16, False), (32, True), (2, False)])
def test_fits_map_layout_and_round_trip(tmp_path, nside, gz):
    from plancklens_b200 import fitsio, hp
    npix = 12 * nside ** 2
    rng = np.random.default_rng(nside)
    fn = str(tmp_path / ('fmask.fits' + ('.gz' if gz else '')))
    m = rng.standard_normal(npix)
    hp.write_map(fn, m)
    if gz:
        assert open(fn, 'rb').read(2) == b'\x1f\x8b' and len(gzip.open(fn).read()) % 2880 == 0
    hdr, d = fitsio.read_hdu(fn, 1)
    rep = 1024 if npix % 1024 == 0 else 1
    assert hdr['TFORM1'] == ('1024D' if rep > 1 else 'D') and hdr['NAXIS2'] == npix // rep
    assert hdr['PIXTYPE'] == 'HEALPIX' and hdr['ORDERING'] == 'RING' and hdr['NSIDE'] == nside
    assert hdr['FIRSTPIX'] == 0 and hdr['LASTPIX'] == npix - 1 and hdr['INDXSCHM'] == 'IMPLICIT'
    assert np.array_equal(hp.read_map(fn), m)
    # three maps, NESTED on disk, read back as RING
    tqu = rng.standard_normal((3, npix))
    hp.write_map(fn, tqu, nest=True)
    q = hp.read_map(fn, field=1)
    assert np.array_equal(q, tqu[1][hp.ring2nest(nside, np.arange(npix))])      # ring[i] = nest[ring2nest(i)]
    t, u = hp.read_map(fn, field=(0, 2), nest=True)
    assert np.array_equal(t, tqu[0]) and np.array_equal(u, tqu[2])


def test_npy_caches_of_earlier_versions_still_load(tmp_path):
    from plancklens_b200 import hp
    alm = rand_alm(np.random.default_rng(1), 12)
    fn = str(tmp_path / 'old_cache.fits')
    with open(fn, 'wb') as f:
        np.save(f, alm)
    assert np.array_equal(hp.read_alm(fn), alm)
    fn2 = str(tmp_path / 'qlm.npy')
    hp.write_alm(fn2, alm)
    assert np.array_equal(np.load(fn2), alm)


def test_sqlite_npdb_layout(tmp_path):
    from plancklens_b200.helpers import sql
    fn = str(tmp_path / 'cldb.db')
    db = sql.npdb(fn)
    v = np.random.default_rng(2).standard_normal(33)
    db.add('sim_qcl_k1p_k2p_lmax32_0000_hash.dat', v)
    assert np.array_equal(db.get('sim_qcl_k1p_k2p_lmax32_0000_hash.dat'), v)
    assert db.get('missing') is None
    db.remove('sim_qcl_k1p_k2p_lmax32_0000_hash.dat')
    assert db.get('sim_qcl_k1p_k2p_lmax32_0000_hash.dat') is None
    # the table layout of the reference (helpers/sql.py:37): id primary key, arr stored as a .npy blob
    con = sqlite3.connect(fn)
    assert con.execute("SELECT sql FROM sqlite_master WHERE name='npdb'").fetchone()[0] == \
        "CREATE TABLE npdb (id STRING PRIMARY KEY, arr ARRAY)"
    db.add('x', v)
    blob = con.execute("SELECT arr FROM npdb WHERE id='x'").fetchone()[0]
    assert bytes(blob[:6]) == b'\x93NUMPY'
    fl = sql.fldb(str(tmp_path / 'fl.db'))
    fl.add('fsky', 0.65)
    assert fl.get('fsky') == 0.65 and fl.get('nope') is None
