"""qtttgym_b200 -- B200-native batched implementation of the game-transition hot path of
Oxel40/qtttgym (quantum tic-tac-toe): reset/step, spooky-mark placement, entanglement-cycle
detection and collapse, the qeval measurement, and uniform-random rollout leaf evaluation.

Export surface mirrors the reference package (qtttgym/__init__.py:1-4: Board, QEvalClassic,
displayBoard, Env) for the parts on the hot path: ``Env`` (single-env adapter), ``BatchedEnv``
(the batched form), ``QEvalB200`` (evaluator plugin), plus ``qeval_both``, ``rollout_eval`` and
``selfplay_sweep``.  Everything computes in libqttt_b200.so (hand-written CUDA for sm_100a);
importing this package without that library works, using it does not.
"""
from .arena import (BatchedStrategy, MCTSStrategy, RandomStrategy, RolloutStrategy, eval_strats,
                    play_games)
from .actions import NUM_ACTIONS, PAIRS, ind2move, move2ind
from .env import (BatchedEnv, Env, get_mask, observe_states, pack_actions, pack_states, render_states, render_text,
                  to_vector, unpack_obs12, unpack_result, unpack_result12)
from .mcts import BatchedMCTS
from .qeval import QEvalB200, qeval_both, square_probabilities
from .vector import VectorEnv
from .rollout import STAT_NAMES, rollout_eval, selfplay_sweep, shard_range, sharded_sweep

__all__ = [
    "NUM_ACTIONS", "PAIRS", "ind2move", "move2ind",
    "BatchedEnv", "Env", "observe_states", "pack_states", "pack_actions", "unpack_result", "unpack_result12", "unpack_obs12",
    "to_vector", "get_mask", "render_text", "render_states",
    "VectorEnv", "BatchedMCTS", "QEvalB200", "qeval_both", "square_probabilities",
    "BatchedStrategy", "MCTSStrategy", "RandomStrategy", "RolloutStrategy", "eval_strats", "play_games",
    "STAT_NAMES", "rollout_eval", "selfplay_sweep", "shard_range", "sharded_sweep",
]
