"""`plancklens` import namespace served by `plancklens_b200`.

Parameter files and drivers written for the reference import `from plancklens.filt import filt_simple, filt_cinv`,
`from plancklens import qest, utils`, `from plancklens.qcinv import opfilt_tt` ... (reference
params/idealized_example.py:26-33, params/anisofilt_example.py, examples/run_qlms.py).  With this directory on
`sys.path` ahead of a reference install those lines resolve, unchanged, to the B200 modules of the same names.

`plancklens.X` IS `plancklens_b200.X` -- the same module object, registered under both names -- so plan caches, the
loaded CUDA library and class identities (`isinstance`, pickled hashes) are shared whichever name a caller used.
`healpy` is left alone: a real healpy keeps working next to this package; where none is installed,
`plancklens_b200.hp.install_as_healpy()` offers the healpy-shaped helper module under that name.
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import sys

import plancklens_b200 as _real

_PREFIX, _REAL = 'plancklens', 'plancklens_b200'


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Resolves `plancklens.a.b` to the already-importable `plancklens_b200.a.b` and hands back that very module."""

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_PREFIX + '.'):
            return None
        real = _REAL + fullname[len(_PREFIX):]
        try:
            rspec = importlib.util.find_spec(real)
        except (ImportError, ValueError):
            return None
        if rspec is None:
            return None
        spec = importlib.machinery.ModuleSpec(fullname, self, origin=rspec.origin,
                                              is_package=rspec.submodule_search_locations is not None)
        spec.has_location = rspec.has_location
        return spec

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_PREFIX):])

    def exec_module(self, module):
        pass   # already executed under its real name


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

# what `os.path.dirname(plancklens.__file__)` is used for: locating data/cls (params/idealized_example.py:38)
__file__ = _real.__file__
__path__ = list(_real.__path__)
__version__ = _real.__version__
