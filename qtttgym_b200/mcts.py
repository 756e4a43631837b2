"""Batched MCTS around the rollout leaf evaluator (reference: mcts.py:132-337).

``BatchedMCTS`` searches one tree per root, all roots concurrently (one thread block per
root; the leaf's ``num_simulations`` playouts run in parallel inside the block).  The API is
the reference's ``Strategy`` shape -- ``reset`` / ``contemplate`` / ``choose`` / ``sync``
(strategy.py:3-36) -- over a batch.  Statistics are bit-identical to the reference's search
when the reference consumes the same keyed random stream (oracle/mcts_oracle.py documents the
keys; tests pin both to recorded reference searches).
"""
from __future__ import annotations

import torch

from . import _lib
from .env import _stream_ptr

ERR_POOL_FULL, ERR_NO_SUCH_CHILD = 1, 2


class BatchedMCTS:
    """mcts.py:132-137: ``MCTS(rollouts=5000, num_simulations=10)``, ``c_puct = 1.0``.

    root_base: global index of root 0 (keys the random stream; lets shards of a larger batch
    reproduce the unsharded search).  max_syncs: how many ``sync`` calls the node pool must
    survive between ``reset``s (a game has at most 9 plies)."""

    def __init__(self, rollouts: int = 5000, num_simulations: int = 10, c_puct: float = 1.0,
                 seed: int = 0, root_base: int = 0, device="cuda", max_syncs: int = 9):
        self.lib = _lib.lib()
        self.num_rollouts, self.num_simulations = int(rollouts), int(num_simulations)
        self.c_puct, self.seed, self.root_base = float(c_puct), int(seed) & 0xFFFFFFFFFFFFFFFF, int(root_base)
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.max_syncs = int(max_syncs)
        self.node_bytes = self.lib.qttt_mcts_node_bytes()
        self.pool = self.meta = None
        self.n_roots = 0

    def reset(self, states, total_rollouts: int | None = None, capacity: int | None = None):
        """mcts.py:139-164 for packed roots int32[R,4].

        The node pool holds ``capacity`` nodes per root.  A rollout adds at most 2 nodes, and
        ``sync`` gives the pruned subtrees back to the pool (``MCTS._prune``, mcts.py:222-231),
        so what has to fit is the LIVE tree: the subtree kept by the last sync plus the rollouts
        since.  Default: ``total_rollouts`` (when given: the whole budget of a search without
        syncs) or three contemplates' worth of rollouts, + 2 per sync.  If a tree ever outgrows
        its pool it stops growing and ``errors()`` reports bit 1; ``live_counts()`` /
        ``peak_counts()`` show the occupancy."""
        self.n_roots = int(states.shape[0])
        if capacity is not None:
            self.capacity = int(capacity)
        else:
            budget = total_rollouts if total_rollouts is not None else 3 * self.num_rollouts
            self.capacity = 1 + 2 * int(budget) + 2 * self.max_syncs
        need = self.n_roots * self.capacity * self.node_bytes
        if self.pool is None or self.pool.numel() < need:
            self.pool = torch.empty(need, dtype=torch.uint8, device=self.device)
        self.meta = torch.zeros((self.n_roots, 8), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_mcts_init(self.pool.data_ptr(), self.capacity, self.meta.data_ptr(),
                                               states.contiguous().data_ptr(), self.n_roots,
                                               _stream_ptr(self.device)))
        return self

    def contemplate(self, n_rollouts: int | None = None):
        """mcts.py:294-306 without the wall clock: exactly ``n_rollouts`` (default
        ``self.num_rollouts``) calls of ``_rollout`` per root."""
        n = self.num_rollouts if n_rollouts is None else int(n_rollouts)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_mcts_run(self.pool.data_ptr(), self.capacity, self.meta.data_ptr(), n,
                                              self.num_simulations, self.c_puct, self.seed, self.root_base,
                                              self.n_roots, _stream_ptr(self.device)))
        return self

    def root_stats(self):
        """(N int32[R,36], Q float64[R,36], Ntot int32[R], choose uint8[R])."""
        r, dev = self.n_roots, self.device
        n = torch.empty((r, 36), dtype=torch.int32, device=dev)
        q = torch.empty((r, 36), dtype=torch.float64, device=dev)
        ntot = torch.empty(r, dtype=torch.int32, device=dev)
        choose = torch.empty(r, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.qttt_mcts_stats(self.pool.data_ptr(), self.capacity, self.meta.data_ptr(),
                                                n.data_ptr(), q.data_ptr(), ntot.data_ptr(), choose.data_ptr(),
                                                r, _stream_ptr(dev)))
        return n, q, ntot, choose

    def choose(self):
        """mcts.py:308-315: uint8[R] best action per root (255 when the root has no action)."""
        return self.root_stats()[3]

    def sync(self, actions, states):
        """mcts.py:317-337: ``actions`` uint8[R] were played and the games are now ``states``."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_mcts_sync(self.pool.data_ptr(), self.capacity, self.meta.data_ptr(),
                                               actions.to(torch.uint8).contiguous().data_ptr(),
                                               states.contiguous().data_ptr(), self.n_roots,
                                               _stream_ptr(self.device)))
        return self

    def errors(self):
        """int32[R] error bits: 1 = node pool exhausted, 2 = sync found no matching child."""
        return self.meta[:, 3]

    def node_counts(self):
        """nodes ever taken from the fresh end of each root's pool (high-water mark)"""
        return self.meta[:, 1]

    def live_counts(self):
        """nodes alive in each tree now (taken minus reclaimed by ``sync``)"""
        return self.meta[:, 1] - self.meta[:, 5]

    def peak_counts(self):
        """most nodes alive at once in each tree since ``reset``"""
        return self.meta[:, 6]
