"""eon_b200 -- B200-native (sm_100a) backend for eOn's batched saddle-search hot path.

Public API (all evaluation happens in libeonb200.so on the GPU; there is no CPU fallback):

* `B200Potential` -- mirror of ``eonc::Potential`` (force / forceBatch / capability queries)
* `Walkers`       -- device-resident batches of saddle searches / relaxations
* `search_params` -- the subset of the reference's Parameters this path reads
* `confile`       -- minimal ``.con`` / ``mode.dat`` reader-writer for fixtures and drivers
"""
from . import _capi, confile, displace, hessian, neb, prefactor, process_search
from ._capi import (MINMODE_DIMER, MINMODE_LANCZOS, OPT_CG, OPT_FIRE, OPT_LBFGS, POT_EAM, POT_LJ, POT_LJCLUSTER,
                    POT_MORSE_PT, ROT_CG, ROT_LBFGS, ROT_SD)
from ._lib import BackendUnavailable
from .potential import B200Potential, PotentialError, eam_row
from .walkers import Harvest, Walkers, search_params

__all__ = ["B200Potential", "PotentialError", "BackendUnavailable", "Walkers", "Harvest", "search_params", "eam_row",
           "confile", "process_search", "displace", "hessian", "prefactor", "neb", "POT_LJ", "POT_LJCLUSTER", "POT_MORSE_PT", "POT_EAM", "OPT_CG", "OPT_LBFGS", "OPT_FIRE",
           "MINMODE_DIMER", "MINMODE_LANCZOS", "ROT_SD", "ROT_CG", "ROT_LBFGS"]
