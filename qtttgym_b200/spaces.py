"""Space descriptors of the env API (reference: qtttgym/env.py:19-25).

The reference declares ``action_space = Tuple((Discrete(9), Discrete(9)))`` and an
``observation_space`` Dict of ``Repeated(Tuple(Discrete(9), Discrete(9)), 5 / 4)``, a ``Box`` for
the classical board and ``Discrete(2)`` for the turn, using gymnasium and
``ray.rllib.utils.spaces.repeated.Repeated``.  Real gymnasium spaces are used here when gymnasium
is importable; otherwise the small stand-ins below provide the same constructor shapes plus
``sample()`` / ``contains()``, so that code touching ``env.action_space`` keeps working.

One correction: the reference declares the classical board as
``Box(-1, 1, shape=(9,))`` although its values are the owning move indices -1..8 (there is a
TODO on that line, env.py:18); the range here is -1..8.
"""
from __future__ import annotations

import random as _random

try:  # pragma: no cover - gymnasium is not installed in the build image
    import gymnasium as _gym
    from gymnasium.spaces import Box, Dict, Discrete, Tuple
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    _gym = None
    HAVE_GYMNASIUM = False

    class _Space:
        def seed(self, seed=None):
            self._rng = _random.Random(seed)
            return [seed]

        @property
        def rng(self):
            if not hasattr(self, "_rng"):
                self._rng = _random.Random()
            return self._rng

        def __contains__(self, x):
            return self.contains(x)

    class Discrete(_Space):
        def __init__(self, n, start=0):
            self.n, self.start = int(n), int(start)

        def sample(self):
            return self.start + self.rng.randrange(self.n)

        def contains(self, x):
            try:
                return int(x) == x and self.start <= int(x) < self.start + self.n
            except (TypeError, ValueError):
                return False

        def __repr__(self):
            return f"Discrete({self.n})"

    class Tuple(_Space):
        def __init__(self, spaces):
            self.spaces = tuple(spaces)

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

        def contains(self, x):
            return hasattr(x, "__len__") and len(x) == len(self.spaces) and \
                all(s.contains(v) for s, v in zip(self.spaces, x))

        def __len__(self):
            return len(self.spaces)

        def __getitem__(self, i):
            return self.spaces[i]

        def __repr__(self):
            return f"Tuple({', '.join(map(repr, self.spaces))})"

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape or ()), dtype

        def sample(self):
            import numpy as np
            return np.array([self.rng.randint(int(self.low), int(self.high)) for _ in range(
                int(np.prod(self.shape)) if self.shape else 1)], dtype=self.dtype or "int32").reshape(self.shape)

        def contains(self, x):
            import numpy as np
            a = np.asarray(x)
            return a.shape == self.shape and bool(((a >= self.low) & (a <= self.high)).all())

        def __repr__(self):
            return f"Box({self.low}, {self.high}, {self.shape})"

    class Dict(_Space):
        def __init__(self, spaces):
            self.spaces = dict(spaces)

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return isinstance(x, dict) and set(x) == set(self.spaces) and \
                all(self.spaces[k].contains(v) for k, v in x.items())

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def __repr__(self):
            return f"Dict({self.spaces})"


class Repeated:
    """``ray.rllib.utils.spaces.repeated.Repeated``: a variable-length list (<= max_len) of a child
    space (ray is not a dependency here, so this stand-in is always used)."""

    def __init__(self, child_space, max_len):
        self.child_space, self.max_len = child_space, int(max_len)

    def sample(self):
        return [self.child_space.sample() for _ in range(_random.randrange(self.max_len + 1))]

    def contains(self, x):
        return isinstance(x, (list, tuple)) and len(x) <= self.max_len and \
            all(self.child_space.contains(v) for v in x)

    def __contains__(self, x):
        return self.contains(x)

    def __repr__(self):
        return f"Repeated({self.child_space!r}, {self.max_len})"


def action_space():
    """env.py:19"""
    return Tuple((Discrete(9), Discrete(9)))


def observation_space():
    """env.py:20-25 (with the classical range corrected to -1..8, quirk Q6).  With real gymnasium
    the two q-state lists cannot be a gymnasium space (Repeated is an RLlib class), so the Dict then
    holds only 'classical' and 'turn' and the Repeated descriptors are exposed beside it."""
    import numpy as np
    pair = Tuple((Discrete(9), Discrete(9)))
    classical = Box(-1, 8, shape=(9,), dtype=np.int32)
    if HAVE_GYMNASIUM:  # pragma: no cover
        return Dict({"classical": classical, "turn": Discrete(2)})
    return Dict({"q_states_p1": Repeated(pair, 5), "q_states_p2": Repeated(pair, 4),
                 "classical": classical, "turn": Discrete(2)})
