"""Batched tournament harness around the transition path (reference: strategy.py:3-36 the
``Strategy`` interface, strat_eval.py:21-95 ``check_win`` / ``play_game`` / ``eval_strats`` and
its (strat1 wins, strat2 wins, draws) tally convention).

A strategy here plays N games at once: ``choose()`` returns one action index per env.  Two
are provided: ``RandomStrategy`` (the uniform-random policy of ``MCTS._simulate``) and
``RolloutStrategy`` (flat Monte-Carlo: every legal action is scored by random rollouts from
both of its collapse outcomes -- ``qeval_both`` + ``rollout_eval`` -- i.e. the leaf evaluator
of mcts.py:166-173 applied one ply deep, without the tree).
"""
from __future__ import annotations

import torch

from .env import BatchedEnv
from .qeval import qeval_both
from .rollout import rollout_eval


class BatchedStrategy:
    """strategy.py:3-36 for a batch: reset / contemplate / choose / sync."""

    def reset(self, env: BatchedEnv):
        self.env = env

    def contemplate(self, thinking_time=None):
        """spend some time planning the move (no-op by default)"""

    def choose(self) -> torch.Tensor:
        """uint8[N] action index per env (255 for envs that are already terminated)"""
        raise NotImplementedError

    def sync(self, actions: torch.Tensor):
        """told which actions were played (stateless strategies ignore it)"""


def _finished(env: BatchedEnv) -> torch.Tensor:
    return env.done


class RandomStrategy(BatchedStrategy):
    """mcts.py:287-292: uniform over the legal actions -- drawn by the step kernel's own random
    policy (``qttt_step_random`` on a scratch copy of the states: Philox keyed (seed, env, ply,
    call)), not by eager tensor ops: one launch, no host synchronisation."""

    def __init__(self, seed: int = 0):
        self.seed = int(seed)
        self.calls = 0
        self._scratch = None

    def reset(self, env):
        super().reset(env)
        self._scratch = BatchedEnv(env.num_envs, device=env.device, seed=self.seed, game_base=env.game_base)

    def choose(self):
        sc = self._scratch
        sc.state.copy_(self.env.state)
        self.calls += 1
        sc.epoch = self.calls                      # a fresh draw per call even at the same ply
        _, _, _, _, info = sc.step_random(record=True)
        return info["action"]                      # 255 for games that are already over


class RolloutStrategy(BatchedStrategy):
    """Flat Monte-Carlo player.  For every env and every legal action: both collapse children
    (mcts.py:233-267), ``n_rollouts`` uniform-random playouts from each (mcts.py:185-208), the
    action's score is the mean child value seen from the mover (children are scored from their
    own side to move, mcts.py:171, hence the sign flip as in ``_backpropogate``, mcts.py:179)."""

    def __init__(self, n_rollouts: int = 32, seed: int = 0):
        self.n_rollouts = int(n_rollouts)
        self.seed = int(seed)
        self.calls = 0

    def choose(self):
        env = self.env
        n, dev = env.num_envs, env.device
        legal = env.action_mask()                                      # bool[N,36]
        states = env.state.unsqueeze(1).expand(n, 36, 4).reshape(n * 36, 4).contiguous()
        acts = torch.arange(36, device=dev, dtype=torch.uint8).repeat(n)
        kids = qeval_both(states, acts, want_boards=False, want_probs=False)
        self.calls += 1
        v0 = rollout_eval(kids["next0"], self.n_rollouts, self.seed + 2 * self.calls)[1]
        v1 = rollout_eval(kids["next1"], self.n_rollouts, self.seed + 2 * self.calls + 1)[1]
        score = (-(v0 + v1) * 0.5).view(n, 36)
        score = torch.where(legal, score, torch.full_like(score, -2.0))
        a = score.argmax(1).to(torch.uint8)
        a[(~legal.any(1)) | _finished(env)] = 255
        return a


class MCTSStrategy(BatchedStrategy):
    """The reference's ``MCTS`` strategy (mcts.py:132-337) for a batch: ``contemplate`` runs
    ``rollouts`` rollouts per live game, ``choose`` returns the best root action, ``sync``
    re-roots every tree on the move actually played (own and opponent's)."""

    def __init__(self, rollouts: int = 200, num_simulations: int = 10, seed: int = 0):
        from .mcts import BatchedMCTS
        self.search = BatchedMCTS(rollouts=rollouts, num_simulations=num_simulations, seed=seed)

    def reset(self, env):
        super().reset(env)
        self.search.device = env.device
        self.search.reset(env.state)
        self._live = ~env.done.clone()

    def contemplate(self, thinking_time=None):
        self.search.contemplate()

    def choose(self):
        a = self.search.choose()
        return torch.where(_finished(self.env), torch.full_like(a, 255), a)

    def sync(self, actions):
        # trees of games that were still running before this move follow it; finished games
        # (action 255) keep their root: qttt_mcts_sync leaves them alone
        self.search.sync(actions, self.env.state)


def play_games(strat_x: BatchedStrategy, strat_o: BatchedStrategy, n_games: int, seed: int = 0,
               device="cuda"):
    """strat_eval.py:34-63 for ``n_games`` games at once: X (player 1) and O alternate until a
    line exists or 9 moves are on the board.  Returns (env, winner uint8[N]: 0 draw, 1 X, 2 O).
    The 9 plies are enqueued back to back: nothing here reads a device value on the host, the
    caller synchronises when it looks at the result (finished games answer 255 = no move)."""
    env = BatchedEnv(n_games, device=device, seed=seed)
    strat_x.reset(env)
    strat_o.reset(env)
    for ply in range(9):
        mover = strat_x if ply % 2 == 0 else strat_o
        mover.contemplate()
        actions = mover.choose()
        env.step(actions)
        strat_x.sync(actions)
        strat_o.sync(actions)
    winner = env.winner()
    return env, winner


def eval_strats(strat1: BatchedStrategy, strat2: BatchedStrategy, num_games: int = 200, seed: int = 0,
                device="cuda"):
    """strat_eval.py:65-95: half of the games with strat1 moving first, half with strat2;
    returns {"strat1_wins", "strat2_wins", "draws", "games"} in the reference's convention."""
    half = num_games // 2
    _, w_a = play_games(strat1, strat2, half, seed, device)
    _, w_b = play_games(strat2, strat1, half, seed + 1, device)
    s1 = int((w_a == 1).sum() + (w_b == 2).sum())
    s2 = int((w_a == 2).sum() + (w_b == 1).sum())
    draws = int((w_a == 0).sum() + (w_b == 0).sum())
    return {"strat1_wins": s1, "strat2_wins": s2, "draws": draws, "games": 2 * half}
