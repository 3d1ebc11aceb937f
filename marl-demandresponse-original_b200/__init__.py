"""B200-native step path of zhimaerfan/marl-demandresponse-original's MADemandResponseEnv.

Import as ``import mdr_b200`` (alias package at the repo root) or
``importlib.import_module("marl-demandresponse-original_b200")``.

Host-only helpers (config flattening, population builders, default config, the ctypes binding)
import without a GPU; the environment classes need torch + a CUDA device and the built
``csrc/libmdr_b200.so`` and raise otherwise -- there is no CPU fallback.
"""
from . import _lib, build, config_flatten, default_config, perlin, population  # noqa: F401
from ._lib import MdrError, load as load_library  # noqa: F401
from .config_flatten import FlatConfig, comm_table  # noqa: F401
from .default_config import default_config as make_default_config  # noqa: F401
from .population import (population_spec, reference_order_population, shard_population,  # noqa: F401
                         synthetic_interp_table, synthetic_population)


def __getattr__(name):
    # torch-dependent classes are imported lazily so that `import mdr_b200` stays cheap
    if name in ("VecDemandResponseEnv", "load_interp_table"):
        from . import vec_env
        return getattr(vec_env, name)
    if name == "MADemandResponseEnv":
        from .env import MADemandResponseEnv
        return MADemandResponseEnv
    if name in ("DeviceRolloutCollector", "ActorMLP"):
        from . import rollout
        return getattr(rollout, name)
    if name in ("regenerate_table", "regenerate_entries"):
        from . import montecarlo
        return getattr(montecarlo, name)
    if name in ("RolloutMetrics", "shard_range", "reduce_metrics", "reduce_device_metrics"):
        from . import sharding
        return getattr(sharding, name)
    raise AttributeError(name)
