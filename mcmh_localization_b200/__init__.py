"""mcmh_localization_b200 -- B200-native Monte-Carlo localization core.

Drop-in for the per-particle filter step of gustavorvillela/mcmh_localization
(app/scripts/parallel_utils.py + the glue arithmetic of app/scripts/amcmh_localizer.py):

* ``Localizer``        stateful, device-resident filter (load_map / set_params / predict /
                       update / estimate / resample), the fast path;
* ``parallel_utils``   function shim with the reference's names and signatures (NumPy in/out).

All arithmetic runs in hand-written sm_100a CUDA kernels inside ``libmcl.so`` (C ABI:
``include/mcl.h``), called through ctypes; PyTorch only owns device buffers and streams.
There is no CPU fallback.
"""
from ._lib import MclError, RESAMPLE_FIXED_POINT, RESAMPLE_REFERENCE_F32  # noqa: F401
from .params import DEFAULT_PARAMS, YAML_PARAMS, load_params  # noqa: F401
from .maps import GridMap, load_map_yaml, map_from_occupancy  # noqa: F401


def __getattr__(name):
    if name == "Localizer":
        from .localizer import Localizer
        return Localizer
    if name == "ShardedLocalizer":
        from .sharded import ShardedLocalizer
        return ShardedLocalizer
    raise AttributeError(name)
