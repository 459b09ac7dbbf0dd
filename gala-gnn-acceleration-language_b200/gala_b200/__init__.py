"""gala_b200 -- B200-native (sm_100a) sparse aggregation kernels behind GALA's
generated-code operator interface.  See DESIGN.md / INTEGRATION.md at the repo root."""
from . import lib, ops, emitted, synth, formats  # noqa: F401
from .ops import TiledGraph  # noqa: F401

__all__ = ["lib", "ops", "emitted", "synth", "formats", "TiledGraph"]
