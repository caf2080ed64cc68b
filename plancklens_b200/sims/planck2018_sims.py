"""Placeholder for the reference's `plancklens.sims.planck2018_sims` (FFP10 simulation readers).

Those classes only read files of the Planck 2018 release from NERSC project directories
(reference sims/planck2018_sims.py:1-22); nothing here can serve them.  The module exists so that the import line of
an unmodified parameter file (`from plancklens.sims import planck2018_sims, phas, maps, utils`, reference
params/idealized_example.py:33) resolves; constructing a library says what to use instead.
"""


class _nersc_only:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "%s reads the FFP10 maps on NERSC, which do not exist here: use cmbs.sims_cmb_unl(cls, phas.lib_phas(...)) "
            "for Gaussian skies with the same spectra" % type(self).__name__)


class cmb_len_ffp10(_nersc_only):
    pass


class cmb_unl_ffp10(_nersc_only):
    pass


class smica_dx12(_nersc_only):
    pass
