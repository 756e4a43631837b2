"""The 36-way action encoding of the reference's search code (mcts.py:339-349; copies in
qttt.py:324-335, alphazero.py:350-361, strat_eval.py:8-19): index k <-> pair (i < j) in
lexicographic order."""
from __future__ import annotations

PAIRS: tuple[tuple[int, int], ...] = tuple((i, j) for i in range(9) for j in range(i + 1, 9))
NUM_ACTIONS = 36


def ind2move(n: int) -> tuple[int, int]:
    """mcts.py:339-343."""
    return PAIRS[int(n)]


def move2ind(i: int, j: int) -> int:
    """mcts.py:345-350 (order of i, j irrelevant)."""
    if i > j:
        i, j = j, i
    return (15 * i - i * i + 2 * j - 2) // 2
