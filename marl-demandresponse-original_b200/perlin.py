"""Host-side grid-signal noise for the dict API (production mode of the drop-in env).

Restates ``utils.Perlin`` (utils.py:1231-1253) over the PyPI ``perlin-noise`` package's 1-D
algorithm.  The package is an un-vendored, unpinned dependency of the reference
(utils.py:8), so the lattice-noise part follows its published algorithm:
``noise(x) = sum_{i in {floor(xo), floor(xo+1)}} fade(1 - |xo - i|) * g_i * (xo - i)`` with
``xo = x * octaves``, ``fade(t) = 6t^5 - 15t^4 + 10t^3`` and
``g_i = Random(seed * max(1, |i + 1|)).uniform(-1, 1)``.
"""
import math
import random as _random


class OctaveNoise:
    def __init__(self, octaves, seed):
        self.octaves, self.seed, self._g = octaves, seed, {}

    def _gradient(self, i):
        g = self._g.get(i)
        if g is None:
            g = _random.Random(self.seed * max(1, int(abs(i + 1)))).uniform(-1, 1)
            self._g[i] = g
        return g

    def noise(self, x):
        xo = x * self.octaves
        total = 0
        for i in (math.floor(xo), math.floor(xo + 1)):
            d = xo - i
            t = 1 - abs(d)
            total += (6 * math.pow(t, 5) - 15 * math.pow(t, 4) + 10 * math.pow(t, 3)) * (self._gradient(i) * d)
        return total


class Perlin:
    """utils.Perlin(amplitude, nb_octaves, octaves_step, period, seed)."""

    def __init__(self, amplitude, nb_octaves, octaves_step, period, seed):
        self.amplitude, self.nb_octaves, self.period = amplitude, nb_octaves, period
        self.noise_list = [OctaveNoise(2 ** i * octaves_step, seed) for i in range(nb_octaves)]

    def calculate_noise(self, x):
        noise = 0
        for j in range(self.nb_octaves - 1):
            noise += self.noise_list[j].noise(x / self.period) / (2 ** j)
        noise += self.noise_list[-1].noise(x / self.period) / (2 ** self.nb_octaves - 1)
        return self.amplitude * noise
