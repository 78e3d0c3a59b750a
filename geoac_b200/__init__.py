"""geoac_b200: B200-native batched ray tracing for GeoAc's `-prop` hot path."""
from . import abi  # noqa: F401
from .api import GeoAcError, Tracer, default_params, load_met_1d, load_met_grid, prop_angles  # noqa: F401
