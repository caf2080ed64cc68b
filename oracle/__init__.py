"""CPU oracle for the plancklens SHT / CG / QE hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED at the SHT seam: the reference (`/root/reference/plancklens/shts.py:4-35`)
delegates every transform to third-party healpy (unpinned, `pyproject.toml:13`) or
lenspyx/ducc0, neither of which is installed here, and the reference's only test
(`tests/test_w.py`) does not touch this path.  The SHT arithmetic below is therefore a
restatement of the published HEALPix / libsharp conventions, checked against closed forms,
sympy/mpmath evaluations of the Wigner-d definition and a brute-force O(npix*nalm) sum.
Everything ABOVE the seam (qcinv operators, cd_solve, multigrid, qest legs) is pinned: the
unmodified reference Python is imported in the build container with `oracle/healpy_shim`
standing in for healpy, and its outputs are committed under `tests/golden/`
(see `tests/golden/make_golden.py`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package.  The product (`plancklens_b200`) never does.
"""
