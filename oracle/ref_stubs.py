"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Stub modules that let the *unmodified* reference (``/root/reference``) be imported in the
build container so that golden vectors can be generated from it (see
``oracle/make_golden.py``).  ``/root/reference`` does not exist on the GPU box, so nothing
under ``tests/ -m gpu``, ``smoke()`` or ``bench.py`` calls :func:`install` at run time.

Missing third-party imports of the reference and what they are used for
(reference file:line):

* ``gym``                      env/MA_DemandResponse.py:3   (dead import)
* ``ray`` + ``ray.rllib...``   env/MA_DemandResponse.py:4,13-15 (base class only)
* ``perlin_noise``             utils.py:8,1243-1245,1251-1252 (grid-signal noise)
* ``matplotlib.pyplot``        utils.py:7 (plotting only)
* ``cvxpy``                    agents/MPC.py:2 (MPC controller, off the hot path)
* ``wandb``                    utils.py (logging only; present in some images)

``perlin_noise`` (PyPI ``perlin-noise``) is NOT vendored by the reference and its version
is pinned nowhere (no requirements file; README lists a different package).  The class
below restates the package's published 1-D algorithm; because the real package is absent,
parity of the perlin value itself is **unpinned** -- the GPU environment is therefore fed
host-replayed perlin values in parity mode (BASELINE.json north_star: "host-replayed
random draws ... perlin signal noise").
"""
from __future__ import annotations

import itertools
import math
import os
import random
import sys
import types

REFERENCE_ROOT = os.environ.get("MDR_REFERENCE_ROOT", "/root/reference")


# --------------------------------------------------------------------------------------
# perlin-noise restatement (published algorithm of PyPI `perlin-noise`, 1.x; unpinned)
# --------------------------------------------------------------------------------------
def _fade(t: float) -> float:
    return 6 * math.pow(t, 5) - 15 * math.pow(t, 4) + 10 * math.pow(t, 3)


def _hasher(coors) -> int:
    return max(
        1,
        int(abs(sum(10 ** i * c for i, c in enumerate(coors)) + 1)),
    )


def _sample_vector(dimensions: int, seed) -> list:
    st = random.getstate()
    random.seed(seed)
    vec = [random.uniform(-1, 1) for _ in range(dimensions)]
    random.setstate(st)
    return vec


class _RandVec:
    def __init__(self, coordinates, seed):
        self.coordinates = coordinates
        self.vec = _sample_vector(len(coordinates), seed)

    def dists_to(self, coordinates):
        return tuple(c1 - c2 for c1, c2 in zip(coordinates, self.coordinates))

    def weight_to(self, coordinates):
        w = 1.0
        for dist in self.dists_to(coordinates):
            w *= _fade(1 - abs(dist))
        return w

    def get_weighted_val(self, coordinates):
        d = self.dists_to(coordinates)
        return self.weight_to(coordinates) * sum(a * b for a, b in zip(self.vec, d))


class PerlinNoise:
    """Value = sum over the 2 lattice neighbours of fade(1-|dx|) * g_i * dx (1-D case)."""

    def __init__(self, octaves=1, seed=None):
        if octaves <= 0:
            raise ValueError("octaves expected to be positive number")
        self.octaves = octaves
        self.seed = seed if seed else random.randint(1, 10 ** 5)
        self.cache = {}

    def __call__(self, coordinates):
        return self.noise(coordinates)

    def noise(self, coordinates):
        if isinstance(coordinates, (int, float)):
            coordinates = [coordinates]
        coordinates = [c * self.octaves for c in coordinates]
        boxes = [(math.floor(c), math.floor(c + 1)) for c in coordinates]
        total = 0
        for coors in itertools.product(*boxes):
            if coors not in self.cache:
                self.cache[coors] = _RandVec(coors, seed=self.seed * _hasher(coors))
            total += self.cache[coors].get_weighted_val(coordinates)
        return total


# --------------------------------------------------------------------------------------
def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "env", "MA_DemandResponse.py"))


def install() -> None:
    """Inject the stubs and put the reference on sys.path (idempotent)."""
    if not reference_available():
        raise RuntimeError(
            "reference tree not found at %s (only present in the build container)" % REFERENCE_ROOT
        )
    os.environ.setdefault("TZ", "UTC")
    try:
        import time as _time

        _time.tzset()
    except Exception:
        pass

    if "gym" not in sys.modules:
        _module("gym")

    class MultiAgentEnv:  # ray.rllib.env.multi_agent_env.MultiAgentEnv stand-in
        def __init__(self, *a, **k):
            pass

    if "ray" not in sys.modules:
        ray = _module("ray")
        rllib = _module("ray.rllib")
        env = _module("ray.rllib.env")
        mae = _module("ray.rllib.env.multi_agent_env", MultiAgentEnv=MultiAgentEnv)
        utils = _module("ray.rllib.utils")
        ann = _module(
            "ray.rllib.utils.annotations", override=lambda c: (lambda f: f), PublicAPI=lambda f: f
        )
        typ = _module("ray.rllib.utils.typing", MultiAgentDict=dict, AgentID=int)
        ray.rllib, rllib.env, rllib.utils = rllib, env, utils
        env.multi_agent_env, utils.annotations, utils.typing = mae, ann, typ

    if "perlin_noise" not in sys.modules:
        _module("perlin_noise", PerlinNoise=PerlinNoise)

    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = _module("matplotlib")
        mpl.pyplot = _module("matplotlib.pyplot")
    for name in ("cvxpy", "wandb"):
        try:
            __import__(name)
        except Exception:
            _module(name)

    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "monteCarlo")):
        if p not in sys.path:
            sys.path.insert(0, p)


def import_reference():
    """Returns (MADemandResponseEnv, normStateDict, config_dict, ref_utils_module)."""
    install()
    from env.MA_DemandResponse import MADemandResponseEnv  # type: ignore
    import utils as ref_utils  # type: ignore
    from config import config_dict  # type: ignore

    return MADemandResponseEnv, ref_utils.normStateDict, config_dict, ref_utils
