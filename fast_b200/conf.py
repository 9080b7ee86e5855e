"""Configuration handling, mirroring the reference's `fast.conf` contract (fast/conf.py:11-116):
same key names and default values, a dict is kept BY REFERENCE and completed in place, a string
must name a `.py` file that defines a dict `p`.  Two keys are new and optional: `RNG` and
`DEVICE` (see DEFAULTS)."""
import importlib.util
import logging

import numpy

logger = logging.getLogger(__name__)

DEFAULTS = {
    # simulation
    'NPXLS': 'auto', 'DX': 'auto', 'NITER': 1000, 'SUBHARM': False, 'FFTW': False,
    'FFTW_THREADS': 1, 'NCHUNKS': 10, 'TEMPORAL': False, 'DT': 0.001, 'LOGFILE': None,
    'LOGLEVEL': 'INFO', 'SEED': None,
    # transmitter / receiver
    'W0': 'opt', 'D_GROUND': 1.0, 'OBSC_GROUND': 0, 'D_SAT': 0.1, 'OBSC_SAT': 0, 'WVL': 1550e-9,
    'AXICON': False, 'POWER': 1, 'SMF': True,
    # turbulence and link
    'H_SAT': 36e6, 'L_SAT': None, 'H_TURB': numpy.array([0, 10e3]),
    'CN2_TURB': numpy.array([100e-15, 100e-15]), 'WIND_SPD': numpy.array([10, 10]),
    'WIND_DIR': numpy.array([90., 0.]), 'L0': numpy.inf, 'l0': 1e-06, 'ZENITH_ANGLE': 0,
    'PROP_DIR': 'up', 'DTHETA': [4, 0], 'TRANSMISSION': 1,
    # adaptive optics
    'AO_MODE': 'AO', 'DSUBAP': 0.02, 'TLOOP': 0.001, 'TEXP': 0.001, 'ALIAS': True, 'NOISE': 0.0,
    'MODAL': False, 'MODAL_MULT': 1, 'ZMAX': None,
    # communications
    'COHERENT': False, 'MODULATION': None, 'EsN0': None,
}

# Keys that exist only in this implementation.  They are looked up with .get(), never written
# into the user's dict, so a reference config round-trips unchanged.
#   RNG    'device' : Philox4x32-10 noise generated inside the CUDA kernel (default)
#          'device-fast' : opt-in cheaper device stream (Philox4x32-7, 40 random bits per complex
#                     sample); statistically equivalent, different realisations
#          'numpy'  : noise drawn on the host from funcs._R in the reference's order
#                     (bit-compatible stream; results match the reference to fp32 accuracy)
#   DEVICE torch device string; default = current CUDA device
#   KEEP_PHS  True: with RNG='numpy', also materialise each chunk's cropped screens in `sim.phs`
#          (slow inspection path, fastb_screens_crop); default False -- screens never reach HBM
EXTRA_DEFAULTS = {'RNG': 'device', 'DEVICE': None, 'KEEP_PHS': False}


class ConfigParser():
    def __init__(self, fname_or_dict):
        if type(fname_or_dict) == dict:
            self.fname = None
            self.config = fname_or_dict
        elif type(fname_or_dict) == str:
            self.fname = fname_or_dict
            self.config = {}
            self.load(fname_or_dict)
        else:
            raise Exception("Either config file name or params dict required")
        self.defaults = {}
        self.set_defaults()
        self.check()

    def load(self, fname):
        """Execute a python config file and take its module-level dict `p`."""
        if fname.rsplit('.', 1)[-1] != 'py':
            raise Exception("Require .py config file")
        spec = importlib.util.spec_from_file_location("", fname)
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        self.config = module.p

    def check(self):
        """Fill every missing key with its default (one warning per key, like the reference)."""
        for key, value in self.defaults.items():
            if key not in self.config:
                logger.warning(f"Config parameter {key} not defined in {self.fname}, "
                               f"setting default value of {value}")
                self.config[key] = value

    def set_defaults(self):
        self.defaults = DEFAULTS
