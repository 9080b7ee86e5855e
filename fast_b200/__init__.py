"""fast_b200: B200-native (sm_100a) implementation of the Monte-Carlo hot path of FAST
(ojdf/fast) behind the reference's `fast.Fast(p).run()` API.

    import fast_b200 as fast
    sim = fast.Fast(params); result = sim.run()

Importing the package requires the in-tree CUDA library (python build_fastb.py)."""
from . import conf, funcs, turbulence_models, ao_power_spectra, dist   # noqa: F401
from . import _lib                                                      # noqa: F401
from .fast import Fast, FastResult, SpatialFrequencies, SpatialFrequencyStruct, load  # noqa: F401
from . import sweep, configs, comms, complete_orbit_simulation           # noqa: F401

__version__ = "0.1.0"
