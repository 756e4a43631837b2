"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's MCTS search (mcts.py) on top
of the oracle game (oracle/qttt_oracle.py), with every random choice drawn from a KEYED
Philox stream so that a batched GPU implementation can reproduce it exactly, plus the shim
that makes the UNMODIFIED reference MCTS consume the same stream (build container only).

Random choices of the reference and their keyed replacements (Philox4x32-10, key = seed):
  * _select:  np.random.choice(node.children[a])            mcts.py:276
      counter (id_lo, id_hi, depth, 2), id = root * 2^32 + rollout;  index = x1 & 1 (two children)
  * _simulate: np.random.choice(node.actions, p=uniform)     mcts.py:193, 287-292
               np.random.choice(nodes)                       mcts.py:195
      the per-ply draw of oracle/qttt_oracle.py policy_draw(seed, id, len(node.moves), 3) with
      id = (root * 2^32 + rollout) * 4096 + sim (one Philox block per two plies):
      action = floor(word * m / 2^32)-th legal action, child index = coin word & 1
  * QEvalClassic.eval's stdlib coin inside MCTS._step (mcts.py:242, 259): forced to 0 then 1,
    so children[a] == [coin-0 outcome, coin-1 outcome].
"""
from __future__ import annotations

import math

from . import qttt_oracle as O

DOMAIN_SELECT = 2
DOMAIN_SIM = 3
MAX_SIMS = 4096


def _draw(seed, ident, ply, domain):
    return O.philox4x32((ident & 0xFFFFFFFF, (ident >> 32) & 0xFFFFFFFF, ply, domain),
                        (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))


class Node:
    """mcts.py:9-37 GameState: a Game plus turn / winner / terminal and per-action statistics."""
    __slots__ = ("game", "turn", "winner", "terminal", "actions", "children", "Ntot", "N", "W", "Q", "P")

    def __init__(self, game: O.Game, turn: bool):
        self.game = game
        self.turn = turn
        w = game.winner()
        self.winner = {0: None, 1: True, 2: False}[w]            # mcts.py:52-65
        self.terminal = w != 0 or len(game.moves) == 9
        self.actions = game.legal_actions()                      # mcts.py:19-27
        self.children = {a: None for a in self.actions}
        self.Ntot = 0
        self.N = {a: 0 for a in self.actions}
        self.W = {a: 0.0 for a in self.actions}
        self.Q = {a: 0.0 for a in self.actions}
        self.P = None


class MCTS:
    """mcts.py:132-337 restated: reset / _rollout / _select / _uct_select / _expand_child /
    _simulate / _backpropogate / choose / sync.  (The transposition dict of the reference is
    keyed by hash(board + moves); `moves` is the full history, so it never merges two distinct
    tree nodes and is not reproduced.)"""

    def __init__(self, rollouts=5000, num_simulations=10, seed=0, root_index=0, c_puct=1.0):
        self.num_rollouts, self.num_simulations = rollouts, num_simulations
        self.seed, self.root_index, self.c_puct = seed, root_index, c_puct
        self.rollout = 0

    def reset(self, game: O.Game):                                # mcts.py:139-164
        self.root = Node(game.clone(), len(game.moves) % 2 == 0)
        self.rollout = 0

    def _step(self, node: Node, action: int):                      # mcts.py:233-267
        a, b = O.PAIRS[action]
        kids = []
        for coin in (0, 1):
            g = node.game.clone()
            collapsed = g.place(a, b, lambda: coin)
            kids.append(Node(g, not node.turn))
            if not collapsed:
                break
        return kids

    def _uct_select(self, node: Node):                             # mcts.py:280-285
        best, best_v = None, None
        for a in node.actions:
            u = node.P[a] * math.sqrt(node.Ntot) / (1 + node.N[a])
            v = node.Q[a] + self.c_puct * u
            if best is None or v > best_v:                         # max(): first maximum wins
                best, best_v = a, v
        return best

    def _select(self, ident):                                      # mcts.py:269-277
        node, path, depth = self.root, [], 0
        while node.P is not None and not node.terminal:
            a = self._uct_select(node)
            if node.children[a] is None:
                node.children[a] = self._step(node, a)             # mcts.py:210-220
            path.append((node, a))
            kids = node.children[a]
            x = _draw(self.seed, ident, depth, DOMAIN_SELECT)
            node = kids[x[1] & 1] if len(kids) == 2 else kids[0]
            depth += 1
        return path, node

    def _simulate(self, leaf: Node, ident):                        # mcts.py:185-208
        if leaf.terminal:
            w = leaf.game.winner()
        else:
            if leaf.P is None:
                leaf.P = {a: 1 / len(leaf.actions) for a in leaf.actions}   # mcts.py:189-191, 287-289
            g = leaf.game.clone()
            while True:
                x = O.policy_draw(self.seed, ident, len(g.moves), DOMAIN_SIM)
                act = O.policy_action(g.legal_mask(), x[0])
                a, b = O.PAIRS[act]
                bit = x[1] & 1
                g.place(a, b, lambda: bit)
                if g.terminal():
                    break
            w = g.winner()
        return {0: 0, 1: 1, 2: -1}[w]

    def _rollout(self):                                            # mcts.py:166-173
        base = (self.root_index << 32) + self.rollout
        path, leaf = self._select(base)
        r_tot = 0
        for s in range(self.num_simulations):
            r = self._simulate(leaf, base * MAX_SIMS + s)
            r_tot += r if leaf.turn else -r
        r = r_tot / self.num_simulations
        for node, a in reversed(path):                             # mcts.py:175-183
            r = -r
            node.W[a] += r
            node.N[a] += 1
            node.Q[a] = node.W[a] / node.N[a]
            node.Ntot += 1
        self.rollout += 1

    def contemplate(self, n_rollouts=None):                        # mcts.py:294-306 (no clock)
        for _ in range(self.num_rollouts if n_rollouts is None else n_rollouts):
            self._rollout()

    def choose(self):                                              # mcts.py:308-315
        best, best_v = None, None
        for a in self.root.actions:
            v = self.root.Q[a] if self.root.N[a] > 0 else -math.inf
            if best is None or v > best_v:
                best, best_v = a, v
        return best

    def sync(self, action, game: O.Game):                          # mcts.py:317-337
        if action not in self.root.children:
            raise Exception("Invalid Action")
        if self.root.children[action] is None:
            self.root.children[action] = self._step(self.root, action)
        for kid in self.root.children[action]:
            if kid.game.board == game.board and kid.game.moves == game.moves:
                self.root = kid
                return
        raise ValueError("the game state is not a child of the root")

    def root_stats(self):
        n = [self.root.N.get(a, 0) for a in range(36)]
        q = [self.root.Q.get(a, 0.0) for a in range(36)]
        return n, q, self.root.Ntot


# ------------------------------------------------------------------ reference shim
class KeyedNumpy:
    """Stands in for the name ``np`` inside the reference's mcts module: ``random.choice``
    draws from the keyed stream above; everything else is numpy."""

    def __init__(self, real_np, seed, root_index):
        self._np = real_np
        self.seed, self.root_index = seed, root_index
        self.rollout = -1
        self.phase = "select"
        self.depth = 0
        self.sim = -1
        self.sim_len = 0
        self.pending = None
        self.random = self            # np.random.choice -> self.choice

    def __getattr__(self, name):
        return getattr(self._np, name)

    # instrumentation hooks (called by the wrappers installed in shim_reference_mcts)
    def begin_rollout(self):
        self.rollout += 1
        self.phase, self.depth, self.sim = "select", 0, -1

    def begin_sim(self, n_moves):
        self.phase = "sim"
        self.sim += 1
        self.sim_len = n_moves
        self.pending = None

    def choice(self, seq, p=None):
        base = (self.root_index << 32) + self.rollout
        if self.phase == "select":
            x = _draw(self.seed, base, self.depth, DOMAIN_SELECT)
            self.depth += 1
            return seq[x[1] & 1] if len(seq) == 2 else seq[0]
        if p is not None:                                          # sample_action, mcts.py:291-292
            x = O.policy_draw(self.seed, base * MAX_SIMS + self.sim, self.sim_len, DOMAIN_SIM)
            self.pending = x
            return seq[(x[0] * len(seq)) >> 32]                    # seq = node.actions, ascending
        x = self.pending                                           # child choice of the same ply
        node = seq[x[1] & 1] if len(seq) == 2 else seq[0]
        self.sim_len = len(node.moves)
        return node


def shim_reference_mcts(ns, seed, root_index, rollouts, num_simulations):
    """Returns an instance of the UNMODIFIED reference MCTS whose random choices follow the keyed
    stream: the module global ``np`` of mcts.py is rebound to a KeyedNumpy, ``_rollout`` and
    ``_simulate`` are wrapped (instrumentation only; the original methods run), and the qeval
    coin inside ``_step`` is forced to 0 then 1."""
    ref = ns.mcts
    fake = KeyedNumpy(ref.np if not isinstance(ref.np, KeyedNumpy) else ref.np._np, seed, root_index)
    ref.np = fake
    mc = ref.MCTS(rollouts=rollouts, num_simulations=num_simulations)
    orig_rollout, orig_sim, orig_step = mc._rollout, mc._simulate, mc._step

    def rollout():
        fake.begin_rollout()
        return orig_rollout()

    def simulate(node):
        fake.begin_sim(len(node.moves))
        return orig_sim(node)

    def step(node, action):
        ns.coin.bits.clear()
        ns.coin.feed(0, 1)
        return orig_step(node, action)

    mc._rollout, mc._simulate, mc._step = rollout, simulate, step
    return mc, fake
