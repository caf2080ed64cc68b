"""numpy arrays and floats in sqlite3 databases (reference: plancklens/helpers/sql.py:14-107): the `npdb` / `fldb`
caches of qecl, qresp and nhl.  Same table layout (`npdb(id, arr ARRAY)` with the array stored as a `.npy` blob,
`fldb(id, fl REAL)`), so databases written by either code open in the other."""
import io
import os
import sqlite3

import numpy as np

from . import mpi


def adapt_array(arr):
    out = io.BytesIO()
    np.save(out, arr)
    out.seek(0)
    return memoryview(out.read())


def convert_array(text):
    out = io.BytesIO(text)
    out.seek(0)
    return np.load(out)


sqlite3.register_adapter(np.ndarray, adapt_array)
sqlite3.register_converter("ARRAY", convert_array)


class _db:
    table, column, coltype = None, None, None

    def __init__(self, fname, idtype="STRING"):
        if not os.path.exists(fname) and mpi.rank == 0:
            con = sqlite3.connect(fname, detect_types=sqlite3.PARSE_DECLTYPES, timeout=3600)
            con.execute("CREATE TABLE %s (id %s PRIMARY KEY, %s %s)" % (self.table, idtype, self.column, self.coltype))
            con.commit()
            con.close()
        mpi.barrier()
        self.con = sqlite3.connect(fname, timeout=3600., detect_types=sqlite3.PARSE_DECLTYPES)

    def _get(self, idx):
        cur = self.con.cursor()
        cur.execute("SELECT %s FROM %s WHERE id=?" % (self.column, self.table), (idx,))
        data = cur.fetchone()
        cur.close()
        return None if data is None else data[0]

    def remove(self, idx):
        if self._get(idx) is None:
            print("%s remove failed!" % self.table)
            return
        self.con.execute("DELETE FROM %s WHERE id=?" % self.table, (idx,))
        self.con.commit()


class npdb(_db):
    """1-D numpy arrays keyed by a string (reference: sql.py:28-66)."""
    table, column, coltype = 'npdb', 'arr', 'ARRAY'

    def add(self, idx, vec):
        if self._get(idx) is not None:
            print("npdb add failed!")
            return
        vec = np.asarray(vec)
        self.con.execute("INSERT INTO npdb (id,  arr) VALUES (?,?)", (idx, vec.reshape((1, len(vec)))))
        self.con.commit()

    def get(self, idx):
        data = self._get(idx)
        return None if data is None else data.flatten()


class fldb(_db):
    """floats keyed by a string (reference: sql.py:68-107)."""
    table, column, coltype = 'fldb', 'fl', 'REAL'

    def add(self, idx, fl):
        if self._get(idx) is not None:
            print("fldb add failed!")
            return
        self.con.execute("INSERT INTO fldb (id,  fl) VALUES (?,?)", (idx, float(fl)))
        self.con.commit()

    def get(self, idx):
        return self._get(idx)
