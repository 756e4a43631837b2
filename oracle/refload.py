"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (Oxel40/qtttgym).

Only usable in the build container, where the reference is mounted read-only at
/root/reference.  It never travels to the GPU box; nothing under ``-m gpu`` tests,
``smoke()`` or ``bench.py`` may import this module at run time.  It is used by

* ``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/``;
* the ``not gpu`` tests that pin ``oracle/qttt_oracle.py`` / ``oracle/qttt_oracle.c``
  against the live reference (skipped when /root/reference is absent).

The reference's ``qtttgym/env.py:5-8`` imports ``gymnasium`` and
``ray.rllib.utils.spaces.repeated`` which are not installed here (and contribute no
arithmetic: a base class plus space descriptors).  We pre-seed ``sys.modules`` with
inert stand-ins so that ``import qtttgym`` / ``import mcts`` work on the untouched files.

The collapse coin (``qeval.py:35``: ``random.choice(entangled_moves[-1][0:2])``) is forced
by rebinding the *name* ``random`` inside the loaded ``qtttgym.qeval`` module namespace to
a ``ForcedCoin`` object.  The reference files themselves are not modified.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("QTTT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "qtttgym", "board.py"))


def _install_stubs() -> None:
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class _Env:  # gymnasium.Env stand-in: only __init__ is ever reached
            def __init__(self, *a, **k):
                pass

        class _Space:
            def __init__(self, *a, **k):
                self.args, self.kwargs = a, k

        spaces = types.ModuleType("gymnasium.spaces")
        for name in ("Discrete", "Tuple", "Dict", "Box"):
            setattr(spaces, name, type(name, (_Space,), {}))
        gym.Env = _Env
        gym.spaces = spaces
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    if "ray.rllib.utils.spaces.repeated" not in sys.modules:
        chain = ["ray", "ray.rllib", "ray.rllib.utils", "ray.rllib.utils.spaces",
                 "ray.rllib.utils.spaces.repeated"]
        prev = None
        for name in chain:
            mod = sys.modules.get(name) or types.ModuleType(name)
            sys.modules[name] = mod
            if prev is not None:
                setattr(prev, name.rsplit(".", 1)[1], mod)
            prev = mod

        class Repeated:
            def __init__(self, *a, **k):
                self.args, self.kwargs = a, k

        prev.Repeated = Repeated


class ForcedCoin:
    """Stands in for the stdlib ``random`` module inside ``qtttgym.qeval``.

    ``choice(seq)`` returns ``seq[bit]`` for the next forced bit; index 0 is the smaller
    square because ``Board.make_move`` normalises moves to ``a < b`` (board.py:16-18).
    """

    def __init__(self):
        self.bits: list[int] = []
        self.used = 0

    def feed(self, *bits: int) -> None:
        self.bits.extend(int(b) & 1 for b in bits)

    def choice(self, seq):
        if not self.bits:
            raise RuntimeError("ForcedCoin: a collapse happened but no coin bit was fed")
        self.used += 1
        return seq[self.bits.pop(0)]


_cache: dict = {}


def load_reference():
    """Returns a namespace with the live reference modules and the coin shim.

    .qtttgym  -- the reference package (Board, QEvalClassic, Env)
    .mcts     -- the reference ``mcts`` module (MCTS, ind2move, move2ind)
    .coin     -- the ForcedCoin bound as ``random`` inside qtttgym.qeval
    """
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import qtttgym  # noqa: the reference package, untouched
    import qtttgym.qeval as ref_qeval
    import mcts as ref_mcts

    coin = ForcedCoin()
    real_random = ref_qeval.random
    ref_qeval.random = coin  # rebinding a module global, not editing the file
    ns = types.SimpleNamespace(qtttgym=qtttgym, mcts=ref_mcts, coin=coin,
                               real_random=real_random, qeval_module=ref_qeval)
    _cache["ns"] = ns
    return ns
