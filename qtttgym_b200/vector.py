"""``gymnasium.vector``-shaped wrapper around ``BatchedEnv``.

``VectorEnv`` follows the vector-env conventions: ``num_envs``, ``single_action_space`` /
``single_observation_space`` (the reference's spaces, qtttgym/env.py:19-25), batched
``action_space`` / ``observation_space`` descriptions, ``reset(seed=, options=)`` ->
``(obs, info)`` and ``step(actions)`` -> ``(obs, rewards, terminations, truncations, infos)`` with
**next-step autoreset** (gymnasium's default ``AutoresetMode.NEXT_STEP``): an env whose episode
ended at step t is reset by the step t+1 call -- its action is ignored and it returns the
first observation of the new episode with reward -0.0 and ``terminated == False``.  The reset
happens inside the step kernel (``QTTT_STEP_AUTORESET_NEXT``): no host round trip, no second
launch.  Observations are device tensors (``classical`` int8[N,9], ``q_states_p1`` int8[N,5,2] and
``q_states_p2`` int8[N,4,2] padded with -1, ``turn`` uint8[N]) decoded by ``qttt_observe``.
"""
from __future__ import annotations

import torch

from . import spaces as _spaces
from .env import BatchedEnv


class VectorEnv:
    metadata = {"autoreset_mode": "next_step", "render_modes": []}

    def __init__(self, num_envs: int, device="cuda", seed: int = 0, game_base: int = 0,
                 reward: str = "reference"):
        """reward: ``"reference"`` -> the reference's ``Env.step`` reward (-1.0 when a line exists,
        else -0.0: quirk Q1); ``"p1"`` -> ``Env._reward()`` (+1 X wins, -1 O wins, 0 otherwise)."""
        if reward not in ("reference", "p1"):
            raise ValueError("reward must be 'reference' or 'p1'")
        self.env = BatchedEnv(num_envs, device=device, seed=seed, game_base=game_base, obs_mode="packed")
        self.num_envs = self.env.num_envs
        self.device = self.env.device
        self.reward_kind = reward
        self.single_action_space = _spaces.action_space()
        self.single_observation_space = _spaces.observation_space()
        self.action_space = _BatchedSpace(self.single_action_space, self.num_envs)
        self.observation_space = _BatchedSpace(self.single_observation_space, self.num_envs)
        self._obs = None
        self.closed = False

    # -- gymnasium.vector API ---------------------------------------------------------------
    def reset(self, *, seed=None, options=None):
        """All envs restart (the seed is ignored like the reference's, env.py:55-57, Q4)."""
        _, info = self.env.reset(seed=seed, options=options)
        return self._observation(), info

    def step(self, actions):
        """actions: uint8[N] action indices 0..35 or int8[N,2] (a, b) pairs (any order)."""
        self._obs, reward, terminated, truncated, info = self.env.step_obs(actions, autoreset="next", out=self._obs)
        if self.reward_kind == "p1":
            reward = info["reward_p1"]
        return self._obs, reward, terminated, truncated, info

    def close(self):
        self.closed = True

    # -- helpers ----------------------------------------------------------------------------
    def _observation(self):
        self._obs = self.env.observation(out=self._obs)
        return self._obs

    def action_masks(self):
        """bool[N,36] legal actions (mcts.py:87-91)."""
        return self.env.action_mask()

    def sample_actions(self):
        """uniform-random legal action index per env, drawn on the device without stepping"""
        legal = self.env.action_mask().float()
        legal[legal.sum(1) == 0, 0] = 1.0
        return torch.multinomial(legal, 1).squeeze(1).to(torch.uint8)

    @property
    def unwrapped(self):
        return self.env


class _BatchedSpace:
    """n copies of a single-env space (what gymnasium.vector.utils.batch_space describes)."""

    def __init__(self, single, n):
        self.single, self.n = single, int(n)

    def sample(self):
        return [self.single.sample() for _ in range(self.n)]

    def __repr__(self):
        return f"Batched({self.single!r}, n={self.n})"
