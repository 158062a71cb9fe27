"""njode_b200 -- B200-native (sm_100a) implementation of the NJ-ODE training / inference hot path
behind the reference's own Python API (HerreraKrachTeichmann/NJODE, NJODE/models.py and the
NJODE/data_utils.py collate contract)."""
from . import models  # noqa: F401
from .models import NJODE  # noqa: F401

__version__ = "0.1.0"
