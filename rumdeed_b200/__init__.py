"""rumdeed_b200 -- B200-native (sm_100a) replacement for RUMDEED's per-timestep hot path.

The product is `librumdeed_b200.so` (CUDA C++, C ABI in include/rumdeed_b200.h); this
package is the thin Python host mirror used by the tests and the benchmark.
"""
from .api import (Config, Counts, Event, HotPath, Rb2Error, StepResult, device_available, load_library,  # noqa: F401
                  planar_config, tip_config)

from .host_api import Simulation  # noqa: F401,E402

__all__ = ["Simulation", "Config", "Counts", "Event", "HotPath", "Rb2Error", "StepResult", "device_available", "load_library",
           "planar_config", "tip_config"]
