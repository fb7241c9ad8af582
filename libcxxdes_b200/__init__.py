"""libcxxdes_b200 -- B200-native ensemble run path for libcxxdes models.

Importing the package does not need a GPU; constructing an `Ensemble` does, and fails loudly
when the in-tree CUDA library (libcxxdes_b200/_lib/libcxxdes_b200.so) is missing.
"""
from . import abi
from .abi import (AlohaParams, ArchSimParams, Clock, HostColumns, MM1Params, MMCParams,
                  MILLISECONDS, MICROSECONDS, NANOSECONDS, PICOSECONDS, SECONDS)
from .program import AllOf, AnyOf, Delay, Program, Until, Wait
from .ensemble import (CxxdesError, Ensemble, EXPORTED_SYMBOLS, device_count, library_path,
                       load_library, plan_kernel)

__all__ = ["abi", "AlohaParams", "ArchSimParams", "Clock", "HostColumns", "MM1Params", "MMCParams",
           "SECONDS", "MILLISECONDS", "MICROSECONDS", "NANOSECONDS", "PICOSECONDS", "CxxdesError",
           "Ensemble", "EXPORTED_SYMBOLS", "device_count", "library_path", "load_library", "plan_kernel", "Program", "Delay", "Until", "AllOf", "AnyOf", "Wait"]
