"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the qtttgym game-transition path.

This module is the *oracle* (checker) for the CUDA path in ``qtttgym_b200``.  It is a
from-scratch restatement, in plain Python, of the algorithm of the reference
(Oxel40/qtttgym); each function cites the reference lines it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product (``qtttgym_b200``) never does.

Parity pinning: the reference ships no tests or golden vectors, so
this oracle is pinned against *outputs of the reference itself run in the build container*:

* ``tests/test_oracle_vs_reference.py`` drives this module and the unmodified reference
  (``oracle/refload.py``) with identical action / coin traces (random legal play, injected
  illegal actions, post-terminal moves) and compares every field after every step;
* ``tests/golden/*.json`` are traces recorded from the live reference by
  ``oracle/make_golden.py`` (committed together with that script); they travel to the GPU
  box where the reference itself is absent;
* the known-answer vectors of SURVEY.md Appendix A.

State of one game (same objects as the reference, board.py:2-7):
    board  : list[int] * 9   -1 = square not classical, else index of the owning move
    moves  : list[(a, b, idx)] with a < b, idx == position; autofill adds (s, s, idx)
    comps  : list[set[int]]  connected components of the graph of *uncollapsed* moves
                             (the reference calls them ``qstructs``)
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Sequence

# --------------------------------------------------------------------------- tables
#: the 8 winning lines, in the order board.py:84-110 scans them
LINES: tuple[tuple[int, int, int], ...] = (
    (0, 1, 2), (3, 4, 5), (6, 7, 8),      # rows      board.py:85-90
    (0, 3, 6), (1, 4, 7), (2, 5, 8),      # columns   board.py:93-98
    (2, 4, 6),                            # anti-diag board.py:101-105
    (0, 4, 8),                            # diag      board.py:106-110
)

#: action index -> (i, j), i < j, lexicographic (mcts.py:339-343 evaluated for n=0..35)
PAIRS: tuple[tuple[int, int], ...] = tuple(
    (i, j) for i in range(9) for j in range(i + 1, 9))
assert len(PAIRS) == 36


def ind2move(n: int) -> tuple[int, int]:
    """mcts.py:339-343 (the float sqrt there evaluates to this table for 0..35)."""
    return PAIRS[n]


def move2ind(i: int, j: int) -> int:
    """mcts.py:345-350."""
    if i > j:
        i, j = j, i
    return (15 * i - i * i + 2 * j - 2) // 2


# --------------------------------------------------------------------------- qeval
def measure(entangled: Sequence[tuple[int, int, int]], coin: int) -> list[int]:
    """Which square every move of a cyclically entangled component collapses into.

    Restates ``QEvalClassic.eval`` (qeval.py:5-51).  ``entangled`` lists the (a, b, idx)
    moves of one component in idx order; the last one closed the cycle.  ``coin`` is the
    index ``random.choice`` picked from ``last[0:2]`` (qeval.py:35): 0 -> the closing move
    falls into its smaller square.
    """
    k = len(entangled)
    where = [-1] * k
    # incidence per square (qeval.py:12-19); positions into ``entangled`` instead of tuples
    touching: list[list[int]] = [[] for _ in range(9)]
    for pos, (a, b, _) in enumerate(entangled):
        touching[a].append(pos)
        touching[b].append(pos)

    # qeval.py:23-31 -- peel pendant squares: a move hanging off the cycle falls into its
    # leaf square; follow the chain inwards while the next square became a leaf itself.
    for start in range(9):
        sq = start
        while len(touching[sq]) == 1:
            pos = touching[sq].pop()
            a, b, _ = entangled[pos]
            inner = b if sq == a else a
            where[pos] = sq
            touching[inner].remove(pos)
            sq = inner

    # qeval.py:35-49 -- the closing move takes the coin; walk the cycle from its larger
    # square back to its smaller square, each move taking the square its predecessor left.
    last = k - 1
    lo, hi = entangled[last][0], entangled[last][1]
    where[last] = (lo, hi)[coin]
    sq = hi
    occupied = where[last] == sq              # ``fell_in_r``
    touching[sq].remove(last)
    while sq != lo:
        pos = touching[sq].pop()
        a, b, _ = entangled[pos]
        other = a if b == sq else b
        # qeval.py:43: move[0] if fell_in_r ^ (move[0]==r) else move[1]
        #   == stay in ``sq`` when it is still free, else go to the far end
        where[pos] = other if occupied else sq
        touching[other].remove(pos)
        occupied = where[pos] == other
        sq = other
    return where


# --------------------------------------------------------------------------- board
class IllegalMove(Exception):
    """board.py:10-15 raises a bare Exception; env.py:41 swallows it."""


class Game:
    """board.py:1-115 restated.  ``coin`` supplies collapse bits (consumed only when a
    collapse happens, exactly once per collapse -- qeval.py:35)."""

    __slots__ = ("board", "moves", "comps", "collapses")

    def __init__(self):
        self.board: list[int] = [-1] * 9          # board.py:5
        self.moves: list[tuple[int, int, int]] = []   # board.py:4
        self.comps: list[set[int]] = []           # board.py:6
        self.collapses = 0                        # bookkeeping only (not in the reference)

    def clone(self) -> "Game":
        g = Game()
        g.board = list(self.board)
        g.moves = list(self.moves)
        g.comps = [set(c) for c in self.comps]
        g.collapses = self.collapses
        return g

    # -- S1 legality, board.py:10-15 (index >= 9 raises IndexError there: also a no-op);
    #    negative indices alias in the reference and are outside the action domain: illegal
    def is_legal(self, a: int, b: int) -> bool:
        if not (0 <= a <= 8 and 0 <= b <= 8):
            return False
        return a != b and self.board[a] == -1 and self.board[b] == -1

    def place(self, a: int, b: int, coin: Callable[[], int]) -> bool:
        """board.py:9-25.  Returns True when the move collapsed a component."""
        if not self.is_legal(a, b):
            raise IllegalMove((a, b))
        if a > b:                                  # board.py:16-18
            a, b = b, a
        self.moves.append((a, b, len(self.moves)))  # board.py:19
        collapsed = self._entangle(a, b, coin)      # board.py:20
        free = [s for s in range(9) if self.board[s] == -1]
        if len(free) == 1:                          # board.py:21-25 autofill
            s = free[0]
            self.board[s] = len(self.moves)
            self.moves.append((s, s, len(self.moves)))
        return collapsed

    def _entangle(self, a: int, b: int, coin: Callable[[], int]) -> bool:
        """board.py:27-69 (update_qstructs)."""
        ia = next((i for i, c in enumerate(self.comps) if a in c), -1)   # board.py:28-33
        ib = next((i for i, c in enumerate(self.comps) if b in c), -2)   # board.py:35-40
        if ia == ib:                                                    # board.py:42 cycle
            comp = self.comps[ia]
            members = [m for m in self.moves if m[0] in comp]           # board.py:44-50
            squares = measure(members, coin())                          # board.py:51
            for (_, _, idx), sq in zip(members, squares):               # board.py:53-54
                self.board[sq] = idx
            self.comps.pop(ib)                                          # board.py:56
            self.collapses += 1
            return True
        if ia >= 0 and ib >= 0:                                         # board.py:58-61
            self.comps[ia] = self.comps[ia] | self.comps[ib]
            self.comps.pop(ib)
        else:                                                           # board.py:62-69
            i = max(ia, ib)
            if i < 0:
                self.comps.append(set())
                i = len(self.comps) - 1
            self.comps[i].update((a, b))
        return False

    # -- S7, board.py:71-115
    def win_rounds(self) -> tuple[int, int]:
        bd = self.board
        px = po = 10
        for l in LINES:
            owners = [bd[s] for s in l]
            if min(owners) < 0:
                continue
            par = [o & 1 for o in owners]
            if par == [0, 0, 0]:
                px = min(px, max(owners))
            elif par == [1, 1, 1]:
                po = min(po, max(owners))
        return (px if px < 10 else -1, po if po < 10 else -1)

    # -- S9 legal mask, mcts.py:19-27 / 87-91 (computed regardless of terminal)
    def legal_mask(self) -> int:
        m = 0
        for k, (i, j) in enumerate(PAIRS):
            if self.board[i] == -1 and self.board[j] == -1:
                m |= 1 << k
        return m

    def legal_actions(self) -> list[int]:
        return [k for k, (i, j) in enumerate(PAIRS)
                if self.board[i] == -1 and self.board[j] == -1]

    # -- S8 winner, mcts.py:52-65 / strat_eval.py:21-32: 1 = X, 2 = O, 0 = none / draw
    def winner(self) -> int:
        px, po = self.win_rounds()
        if px > 0 and po > 0:
            return 1 if px < po else 2
        if px > 0:
            return 1
        if po > 0:
            return 2
        return 0

    def terminal(self) -> bool:
        """mcts.py:52-65: a line exists or 9 entries in ``moves``."""
        return self.winner() != 0 or len(self.moves) == 9

    # -- env.py:87-112 (_reward: p1 perspective, earlier round wins)
    def reward_p1(self) -> float:
        px, po = self.win_rounds()
        px = 10 if px < 0 else px
        po = 10 if po < 0 else po
        if px < po:
            return 1.0
        if po < px:
            return -1.0
        return 0.0

    # -- env.py:68-85
    def observation(self) -> dict:
        classical = set(self.board)
        q1 = [(a, b) for (a, b, i) in self.moves if i not in classical and i % 2 == 0]
        q2 = [(a, b) for (a, b, i) in self.moves if i not in classical and i % 2 == 1]
        return {"q_states_p1": q1, "q_states_p2": q2,
                "classical": list(self.board), "turn": len(self.moves) % 2}


def coin_from(bits: Iterable[int]) -> Callable[[], int]:
    it: Iterator[int] = iter(bits)
    return lambda: int(next(it)) & 1


class Env:
    """env.py:15-66 restated (gym-like single env).  ``coin`` as in ``Game.place``."""

    def __init__(self, coin: Callable[[], int] | None = None):
        self.game = Game()
        self.coin = coin
        self.last_status = 0

    def reset(self, *, seed=None, options=None):
        """env.py:55-57: seed / options are ignored (Q4)."""
        self.game = Game()
        return self.game.observation(), {}

    def turn(self) -> int:
        return len(self.game.moves)            # env.py:65-66

    def step(self, action, coin: Callable[[], int] | None = None):
        """env.py:34-53.  Illegal action -> state unchanged (Q2); reward is the
        reference's ``(-1 ** cur_player) * float(win)`` == -1.0 * win, i.e. -0.0 / -1.0 (Q1)."""
        try:
            self.game.place(int(action[0]), int(action[1]), coin or self.coin)
            self.last_status = 0
        except IllegalMove:
            self.last_status = 1
        obs = self.game.observation()
        px, po = self.game.win_rounds()
        win = px > 0 or po > 0
        r = -1.0 * float(win)                  # env.py:49 (parses as -(1**p) * float)
        terminated = win or self.turn() > 8     # env.py:51
        return obs, r, terminated, False, {}


# --------------------------------------------------------------------------- Philox
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_U32 = 0xFFFFFFFF


def philox4x32(ctr: Sequence[int], key: Sequence[int], rounds: int = 10) -> tuple[int, int, int, int]:
    """Philox4x32-R (Salmon et al., SC'11; Random123 v1.x).  Not part of the reference:
    the random-policy modes of the CUDA path define their random stream with it, and the
    oracle replays the same stream to check those modes."""
    c0, c1, c2, c3 = (int(x) & _U32 for x in ctr)
    k0, k1 = (int(x) & _U32 for x in key)
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _U32, p1 & _U32, \
                         ((p0 >> 32) ^ c3 ^ k1) & _U32, p0 & _U32
        k0 = (k0 + _W0) & _U32
        k1 = (k1 + _W1) & _U32
    return c0, c1, c2, c3


DOMAIN_STEP = 0      # env.step random policy and the self-play sweep
DOMAIN_ROLLOUT = 1   # rollout leaf evaluation


def policy_draw(seed: int, game_id: int, ply: int, domain: int) -> tuple[int, int]:
    """The random draw of one ply, (action word, coin word): one Philox4x32-10 block with
    counter (game_lo, game_hi, ply >> 1, domain) and key = seed serves two consecutive plies --
    the even ply takes (x0, x1), the odd ply (x2, x3).  Bit 0 of the coin word is the coin."""
    x = philox4x32((game_id & _U32, (game_id >> 32) & _U32, ply >> 1, domain),
                   (seed & _U32, (seed >> 32) & _U32))
    return (x[2], x[3]) if ply & 1 else (x[0], x[1])


def nth_set_bit(mask: int, k: int) -> int:
    for pos in range(64):
        if (mask >> pos) & 1:
            if k == 0:
                return pos
            k -= 1
    raise ValueError("k out of range")


def policy_action(mask: int, x0: int) -> int:
    """uniform legal action: the floor(x0 * m / 2^32)-th set bit of the 36-bit legal mask."""
    m = bin(mask).count("1")
    return nth_set_bit(mask, (x0 * m) >> 32)


def random_playout(game: Game, seed: int, game_id: int, domain: int,
                   trace: list | None = None) -> tuple[int, int, int]:
    """mcts.py:185-208 (_simulate) with the Philox policy: until terminal, action ~ U(legal)
    (mcts.py:287-292), collapse ~ U{0,1} (mcts.py:195 picks uniformly between the two
    children _step enumerates).  Mutates ``game``.  Returns (winner, env_steps, collapses)."""
    steps = collapses = 0
    while not game.terminal():
        ply = len(game.moves)
        x0, x1 = policy_draw(seed, game_id, ply, domain)
        act = policy_action(game.legal_mask(), x0)
        a, b = PAIRS[act]
        bit = x1 & 1
        did = game.place(a, b, lambda: bit)
        if trace is not None:
            trace.append((act, bit))
        steps += 1
        collapses += int(did)
    return game.winner(), steps, collapses


def plies(moves: Sequence[tuple[int, int, int]]) -> int:
    """Number of make_move calls that produced ``moves``: the autofill entry (s, s, idx)
    (board.py:25) is not a ply of its own."""
    return len(moves) - int(bool(moves) and moves[-1][0] == moves[-1][1])


def leaf_value(tally: Sequence[int], n_plies: int) -> float:
    """mcts.py:166-173: mean of r (+1 X / -1 O / 0) taken from the leaf's side to move.
    ``turn`` starts True on the empty board (mcts.py:141) and is flipped once per ply
    (mcts.py:243, 264), so ``leaf.turn == (n_plies % 2 == 0)`` -- autofill does not flip it."""
    xw, ow, dr = tally
    tot = xw + ow + dr
    r = (xw - ow) / tot
    return r if n_plies % 2 == 0 else -r
