"""Build-container only: the oracle against the LIVE unmodified reference (skipped elsewhere)."""
import random

import pytest

from oracle import tracegen as T
from oracle.refload import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not mounted")


@pytest.mark.parametrize("seed,illegal,overrun", [(1, 0.0, 0), (2, 0.15, 0), (3, 0.1, 4)])
def test_python_oracle_equals_live_reference(seed, illegal, overrun):
    ns = load_reference()
    rng = random.Random(seed)
    for _ in range(1500):
        trace = T.random_trace(rng, illegal_rate=illegal, overrun=overrun)
        assert T.replay_oracle(trace) == T.replay_reference(trace, ns)


def test_ind2move_table_matches_reference():
    from oracle import qttt_oracle as O
    ns = load_reference()
    for n in range(36):
        assert ns.mcts.ind2move(n) == O.ind2move(n)
        assert ns.mcts.move2ind(*O.PAIRS[n]) == n == O.move2ind(*O.PAIRS[n])
