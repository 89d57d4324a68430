"""gridcount_b200 -- B200-native (sm_100a) implementation of gridcount's hot path.

The product is the C-ABI shared library `libgridcount_b200.so`
(include/gridcount_b200.h) plus the C host tools under host/.  This Python
package is a thin ctypes mirror of that ABI used by the tests and bench.py; it
performs no computation of its own and has no CPU fallback."""
from . import _cabi
from .api import (GridCounter, Geometry, analyse, gridcalc, setup_geometry, frame_weights, header_string, grid_write,
                  grid_read, xdr_encode, DeviceFrames, PinnedFrames, XtcBuffer, make_gen, device_count, selftest_l2, pack_frames)

__all__ = ["GridCounter", "Geometry", "analyse", "gridcalc", "setup_geometry", "frame_weights", "header_string",
           "grid_write", "grid_read", "xdr_encode", "DeviceFrames", "PinnedFrames", "XtcBuffer", "make_gen", "device_count", "selftest_l2", "pack_frames",
           "_cabi"]
