"""stereo-vision_b200: B200-native (sm_100a) dense-stereo hot path behind libelas's Elas::process.

The product is the C-ABI library libelas_b200.so (include/elas_b200.h) plus the C++ drop-in
`Elas` class under dropin/.  This Python module is only the ctypes view of that ABI used by the
parity tests, bench.py and __graft_entry__.py; it contains no algorithmic code and no CPU fallback.
"""
from .elas_b200 import (ElasB200, Params, LibraryMissing, build_library, load_library,  # noqa: F401
                        robotics, middlebury, stereomapper, demo, LIB_PATH)
from . import synth  # noqa: F401
