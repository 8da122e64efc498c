"""rayoptics_b200 -- a B200 (sm_100a) sequential real-ray trace engine that
drops into mjhoptics/ray-optics models: same ``trace()/trace_raw()`` call
surface, hand-written CUDA underneath (include/b200rt.h)."""
from . import model, roa          # noqa: F401  (roa registers its media classes)

__version__ = '0.1.0'
