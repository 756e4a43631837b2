"""Random-rollout leaf evaluation and the self-play sweep (reference: mcts.py:166-173
``_rollout`` averaging, 185-208 ``_simulate`` / ``_reward``; tally convention of
strat_eval.py:66-94)."""
from __future__ import annotations

import torch

from . import _lib
from .env import _stream_ptr

STAT_NAMES = ("x_wins", "o_wins", "draws", "env_steps", "collapses", "games")


def rollout_eval(roots, n_rollouts: int, seed: int = 0, out=None):
    """``n_rollouts`` uniform-random playouts from every packed root (int32[R,4]).

    Returns ``(tallies int32[R,3] = (X wins, O wins, draws), value f32[R], env_steps int)``;
    ``value`` is the mean playout reward from the root's side to move, exactly the quantity
    ``MCTS._rollout`` backs up (mcts.py:168-173).  Playout j of root r draws from Philox with
    game id ``r * n_rollouts + j`` in domain 1.  ``out``: a tuple returned by an earlier call
    with the same number of roots, reused instead of allocating (``env_steps`` is zeroed).
    """
    lib = _lib.lib()
    dev, r = roots.device, roots.shape[0]
    if out is not None:
        tallies, value, steps = out
        steps.zero_()
    else:
        tallies = torch.empty((r, 3), dtype=torch.int32, device=dev)
        value = torch.empty(r, dtype=torch.float32, device=dev)
        steps = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.qttt_rollout(roots.data_ptr(), r, int(n_rollouts),
                                    int(seed) & 0xFFFFFFFFFFFFFFFF, tallies.data_ptr(),
                                    value.data_ptr(), steps.data_ptr(), _stream_ptr(dev)))
    return tallies, value, steps


def selfplay_sweep(game_lo: int, game_hi: int, seed: int = 0, device="cuda", out=None):
    """Plays games ``[game_lo, game_hi)`` from the empty board to termination with the random
    policy, on ``device``.  Returns int64[16]: the six ``STAT_NAMES`` then the histogram of
    games by number of env-steps (0..9).  Adds into ``out`` when given (stream-ordered, no sync).
    """
    lib = _lib.lib()
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    stats = out if out is not None else torch.zeros(16, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.qttt_sweep(int(game_lo), int(game_hi), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                  stats.data_ptr(), _stream_ptr(dev)))
    return stats


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split of ``[0, total)`` (rank r of W takes ``[r*G/W, (r+1)*G/W)``)."""
    return (total * rank) // world, (total * (rank + 1)) // world


def sharded_sweep(total_games: int, seed: int = 0, device="cuda", group=None, sweep_fn=None):
    """Config 5: every rank plays its slice of the global game-id range, then ONE
    ``all_reduce(SUM)`` of the int64[16] tallies (NCCL over NVLink when the process group is
    nccl).  RNG is keyed on the global game id, so the result is identical for any world
    size.  ``sweep_fn(lo, hi, seed, device) -> int64[16]`` defaults to ``selfplay_sweep`` (the
    CUDA kernel); the CPU tests of this host logic inject a stand-in and a gloo group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_range(total_games, rank, world)
    stats = (sweep_fn or selfplay_sweep)(lo, hi, seed, device)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
