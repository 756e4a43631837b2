"""Backend-agnostic parity checks: every function takes a backend from tests/backends.py and
diffs it against the oracle (oracle/c_oracle.py, pinned to the live reference) or against the
fixtures recorded from the reference (tests/golden).  The GPU tests run them on the CUDA
path; the CPU tests run the same checks on the host emulation of the same source."""
from __future__ import annotations

import numpy as np

from oracle import c_oracle as CO
from oracle import qttt_oracle as O

from helpers import load_golden, run_traces_and_check

PAIRS = np.array(O.PAIRS, dtype=np.int8)
OBS_KEYS = ("classical", "moves", "n_moves", "q_p1", "q_p2", "turn", "rounds", "reward_p1", "winner")


def expand_mask(mask_u64):
    bits = np.arange(36, dtype=np.uint64)
    return ((mask_u64[:, None] >> bits) & np.uint64(1)).astype(bool)


def assert_same_step(out, ref, where):
    assert np.array_equal(out["reward"].view(np.uint32), ref["reward"].view(np.uint32)), f"{where}: reward bits"
    assert np.array_equal(out["done"], ref["done"]), f"{where}: done"
    assert np.array_equal(out["mask"], ref["mask"]), f"{where}: legal mask"
    assert np.array_equal(out["status"], ref["status"]), f"{where}: status"


def assert_same_obs(obs, ref, where):
    for k in OBS_KEYS:
        assert np.array_equal(obs[k], ref[k]), f"{where}: {k}"
    if "mask_bool" in obs:
        assert np.array_equal(obs["mask_bool"].astype(bool), expand_mask(_mask_from_obs(ref))), f"{where}: mask_bool"


def _mask_from_obs(ref):
    free = ref["classical"] < 0
    m = np.zeros(free.shape[0], np.uint64)
    for k, (i, j) in enumerate(O.PAIRS):
        m |= (free[:, i] & free[:, j]).astype(np.uint64) << np.uint64(k)
    return m


# ------------------------------------------------------------------ golden fixtures
def check_golden_traces(backend):
    for fixture in ("kat_appendix_a.json.gz", "traces_v1.json.gz"):
        data = load_golden(fixture)
        games = list(data.values()) if isinstance(data, dict) else data
        run_traces_and_check(backend.games, [g["trace"] for g in games],
                             [g["records"] for g in games], where=f"{backend.name}:{fixture}")


def check_golden_qeval(backend):
    cases = load_golden("qeval_v1.json.gz")
    n = len(cases)
    classical = np.full((n, 9), -1, np.int8)
    moves = np.full((n, 9, 2), -1, np.int8)
    n_moves = np.zeros(n, np.uint8)
    acts = np.zeros(n, np.uint8)
    for i, case in enumerate(cases):
        ent = case["entangled"]
        for r, (a, b, _) in enumerate(ent[:-1]):
            moves[i, r] = (a, b)
        n_moves[i] = len(ent) - 1
        acts[i] = O.move2ind(ent[-1][0], ent[-1][1])
    res = backend.games(n).load(classical, moves, n_moves).qeval_both(acts)
    for i, case in enumerate(cases):
        k = len(case["entangled"])
        assert res["closes"][i] == 1
        assert res["sq0"][i][:k].tolist() == case["out0"], i
        assert res["sq1"][i][:k].tolist() == case["out1"], i
        assert (res["sq0"][i][k:] == -1).all() and (res["sq1"][i][k:] == -1).all()


def check_golden_mcts_step(backend):
    """mcts.py:233-267 children (both collapse outcomes), action lists, winner / terminal."""
    recs = load_golden("mcts_step_v1.json.gz")
    n = len(recs)
    classical = np.array([r["board"] for r in recs], np.int8)
    moves = np.full((n, 9, 2), -1, np.int8)
    n_moves = np.zeros(n, np.uint8)
    for i, r in enumerate(recs):
        for m in r["moves"]:
            moves[i, m[2]] = m[:2]
        n_moves[i] = len(r["moves"])
    games = backend.games(n).load(classical, moves, n_moves)
    obs = games.observe()
    for i, r in enumerate(recs):
        assert obs["mask_bool"][i].astype(bool).tolist() == r["mask"]
    acts = np.array([r["action"] for r in recs], np.uint8)
    res = games.qeval_both(acts, squares=False)      # what MCTS expansion uses: the one-sweep path
    kids = [games.with_state(res["next0"]).observe(), games.with_state(res["next1"]).observe()]
    for i, r in enumerate(recs):
        assert int(res["closes"][i]) == (len(r["children"]) == 2), i
        for c, ch in enumerate(r["children"]):
            o = kids[c]
            assert o["classical"][i].tolist() == ch["board"], (i, c)
            nm = int(o["n_moves"][i])
            assert [m + [j] for j, m in enumerate(o["moves"][i][:nm].tolist())] == ch["moves"], (i, c)
            assert {0: None, 1: True, 2: False}[int(o["winner"][i])] == ch["winner"]
            assert (int(o["winner"][i]) != 0 or nm == 9) == ch["terminal"]
            assert np.nonzero(o["mask_bool"][i])[0].tolist() == ch["actions"]


# ------------------------------------------------------------------ differential vs C oracle
def random_actions(rng, mask_u64, illegal_rate):
    """uniform legal (a, b) per env (random order), with injected illegal pairs."""
    n = mask_u64.shape[0]
    legal = expand_mask(mask_u64)
    score = rng.random((n, 36)) * legal
    k = score.argmax(1)
    pairs = PAIRS[k].copy()
    none = ~legal.any(1)
    pairs[none] = (-1, -1)
    flip = rng.random(n) < 0.5
    pairs[flip] = pairs[flip][:, ::-1]
    if illegal_rate > 0:
        bad = rng.random(n) < illegal_rate
        kind = rng.integers(0, 4, n)
        r1 = rng.integers(0, 9, n).astype(np.int8)
        r2 = rng.integers(0, 9, n).astype(np.int8)
        big = rng.integers(9, 128, n).astype(np.int8)
        neg = rng.integers(-128, 0, n).astype(np.int8)
        cand = np.stack([
            np.stack([r1, r1], 1),                 # same square
            np.stack([big, r2], 1),                # off the board
            np.stack([r1, neg], 1),                # negative index (outside the action domain)
            np.stack([r1, r2], 1),                 # arbitrary pair: often hits a classical square
        ], 0)
        pairs[bad] = cand[kind[bad], np.nonzero(bad)[0]]
    return pairs


def check_random_play(backend, n_games, seed, illegal_rate=0.0, overrun=False, fmt="pair",
                      full_obs_every=1):
    """config 2: n_games envs stepped ply by ply on identical (action, coin) traces; every
    output compared bit for bit after every call."""
    rng = np.random.default_rng(seed)
    ref = CO.Games(n_games)
    dut = backend.games(n_games)
    mask = np.full(n_games, (1 << 36) - 1, np.uint64)
    done = np.zeros(n_games, bool)
    total_steps = 0
    for ply in range(14 if (illegal_rate > 0 or overrun) else 9):
        pairs = random_actions(rng, mask, illegal_rate)
        if not overrun:
            pairs[done] = (-1, -1)             # finished envs idle (illegal no-op)
        coins = rng.integers(0, 2, n_games).astype(np.uint8)
        r = ref.step(pairs, coins)
        if fmt == "pair":
            o = dut.step(pairs, coins)
        else:
            a, b = pairs[:, 0].astype(np.int16), pairs[:, 1].astype(np.int16)
            ok = (a >= 0) & (a < 9) & (b >= 0) & (b < 9) & (a != b)
            lo, hi = np.minimum(a, b), np.maximum(a, b)
            idx = np.where(ok, (15 * lo - lo * lo + 2 * hi - 2) // 2, rng.integers(36, 256, n_games))
            o = dut.step_index(idx.astype(np.uint8), coins)
        where = f"{backend.name} seed {seed} ply {ply}"
        assert_same_step(o, r, where)
        if ply % full_obs_every == 0 or ply >= 8:
            assert_same_obs(dut.observe(), ref.observe(), where)
        total_steps += int((r["status"] == 0).sum())
        mask, done = r["mask"], r["done"].astype(bool)
    assert_same_obs(dut.observe(), ref.observe(), f"{backend.name} seed {seed} final")
    return total_steps


def check_pack_observe_roundtrip(backend, n_games=2000, seed=5):
    rng = np.random.default_rng(seed)
    ref = CO.Games(n_games)
    dut = backend.games(n_games)
    mask = np.full(n_games, (1 << 36) - 1, np.uint64)
    for ply in range(9):
        stop = rng.random(n_games) < 0.15      # freeze some games at every depth
        pairs = random_actions(rng, mask, 0.0)
        pairs[stop] = (-1, -1)
        coins = rng.integers(0, 2, n_games).astype(np.uint8)
        mask = ref.step(pairs, coins)["mask"]
        dut.step(pairs, coins)
    obs = ref.observe()
    packed = backend.games(n_games).load(obs["classical"], obs["moves"], obs["n_moves"])
    assert np.array_equal(np.asarray(packed.state), np.asarray(dut.state)), "pack(observe(s)) != s"
    assert_same_obs(packed.observe(), obs, "roundtrip")


def harvest_positions(n_target, seed, want="closing"):
    """Positions from random self-play on the oracle.
    want='closing': (position, action) pairs at the moment a cycle-closing action is chosen;
    want='any': positions at a uniformly random ply with a random legal action."""
    rng = np.random.default_rng(seed)
    out_cl, out_mv, out_nm, out_act = [], [], [], []
    got = 0
    while got < n_target:
        n = 4096
        ref = CO.Games(n)
        mask = np.full(n, (1 << 36) - 1, np.uint64)
        alive = np.ones(n, bool)
        for ply in range(9):
            legal = expand_mask(mask)
            k = (rng.random((n, 36)) * legal).argmax(1).astype(np.uint8)
            coins = rng.integers(0, 2, n).astype(np.uint8)
            before = ref.observe()
            both = ref.qeval_both(k)
            if want == "closing":
                take = alive & legal.any(1) & (both["closes"] == 1)
            else:
                take = alive & legal.any(1) & (rng.random(n) < 0.15)
            if take.any():
                out_cl.append(before["classical"][take]); out_mv.append(before["moves"][take])
                out_nm.append(before["n_moves"][take]); out_act.append(k[take])
                got += int(take.sum())
            pairs = PAIRS[k].copy()
            pairs[~alive | ~legal.any(1)] = (-1, -1)
            r = ref.step(pairs, coins)
            mask = r["mask"]
            alive &= ~r["done"].astype(bool)
    cat = lambda xs: np.concatenate(xs)[:n_target]   # noqa: E731
    return cat(out_cl), cat(out_mv), cat(out_nm), cat(out_act)


def check_qeval_both(backend, n_boards, seed):
    """config 3: both outcomes bit-exact vs two forced-coin oracle calls; probabilities within
    1e-6 (they are exactly 0, 1/2 or 1)."""
    for want in ("closing", "any"):
        cl, mv, nm, act = harvest_positions(n_boards, seed, want)
        ref = CO.Games.from_arrays(cl, mv, nm)
        want_res = ref.qeval_both(act)
        dut = backend.games(len(act)).load(cl, mv, nm)
        res = dut.qeval_both(act)
        for k in ("closes", "sq0", "sq1"):
            assert np.array_equal(res[k], want_res[k]), f"{backend.name} qeval {want}: {k}"
        # the one-sweep path (no per-move squares: both rootings grown in one loop) must give the
        # very same successors, boards, closes and probabilities as the per-coin path
        fast = dut.qeval_both(act, squares=False)
        for k in ("closes", "next0", "next1", "board0", "board1", "result_prob"):
            assert np.array_equal(fast[k], res[k]), f"{backend.name} qeval {want} one-sweep: {k}"
        lean = dut.qeval_both(act, boards_only=True)      # boards + closes only (config 3's shape)
        for k in ("closes", "board0", "board1"):
            assert np.array_equal(lean[k], res[k]), f"{backend.name} qeval {want} boards-only: {k}"
        assert np.array_equal(res["board0"], want_res["out0"]) and np.array_equal(res["board1"], want_res["out1"])
        probs = np.zeros((len(act), 3), np.float64)
        for c in (0, 1):
            branch = ref.copy()
            branch.step(PAIRS[act], np.full(len(act), c, np.uint8))
            bo = branch.observe()
            assert_same_obs(dut.with_state(res[f"next{c}"]).observe(), bo, f"{backend.name} qeval next{c}")
            probs[:, 0] += 0.5 * (bo["winner"] == 1)
            probs[:, 1] += 0.5 * (bo["winner"] == 2)
        probs[:, 2] = 1.0 - probs[:, 0] - probs[:, 1]
        assert np.abs(res["result_prob"].astype(np.float64) - probs).max() <= 1e-6
        if want == "closing":
            assert (res["closes"] == 1).all()
            assert (res["board0"] != res["board1"]).all()      # the two outcomes always differ


def check_rollout(backend, n_roots, n_rollouts, seed):
    """config 4: per-root tallies identical to the oracle playing the same Philox stream from
    the same roots; value per mcts.py:171-173."""
    cl, mv, nm, _ = harvest_positions(n_roots, seed, "any")
    # add the empty root, a terminal root and an autofilled root when available
    ref = CO.Games.from_arrays(cl, mv, nm)
    want_t, want_steps = ref.rollout(n_rollouts, seed)
    dut = backend.games(len(nm)).load(cl, mv, nm)
    tallies, value, steps = dut.rollout(n_rollouts, seed)
    assert np.array_equal(tallies, want_t), f"{backend.name} rollout tallies"
    assert steps == want_steps
    assert (tallies.sum(1) == n_rollouts).all()
    autofill = (nm == 9) & (mv[:, 8, 0] == mv[:, 8, 1])
    plies = nm.astype(np.int64) - autofill
    r = (tallies[:, 0] - tallies[:, 1]).astype(np.float32) / np.float32(n_rollouts)
    want_v = np.where(plies % 2 == 0, r, -r).astype(np.float32)
    assert np.abs(value - want_v).max() <= 1e-6


def check_rollout_terminal_roots(backend):
    """terminal and autofilled roots: zero-length playouts, value sign from plies (mcts.py:243)."""
    kat = load_golden("kat_appendix_a.json.gz")
    recs = [kat["kat8"]["records"][-1], kat["kat7"]["records"][-1], kat["kat10"]["records"][-1]]
    n = len(recs)
    classical = np.array([r["board"] for r in recs], np.int8)
    moves = np.full((n, 9, 2), -1, np.int8)
    nmv = np.array([len(r["moves"]) for r in recs], np.uint8)
    for i, r in enumerate(recs):
        for m in r["moves"]:
            moves[i, m[2]] = m[:2]
    tallies, value, steps = backend.games(n).load(classical, moves, nmv).rollout(8, 3)
    assert steps == 0
    assert tallies.tolist() == [[0, 8, 0], [8, 0, 0], [0, 8, 0]]
    # kat8: 9 entries, 8 plies -> turn True -> value = r = -1; kat7: 5 plies -> -(+1); kat10: 7 plies -> -(-1)
    assert value.tolist() == [-1.0, -1.0, 1.0]


def check_sweep(backend, n_games, seed):
    """config 5: tallies identical to the oracle's self-play on the same Philox stream, and
    additive over shards of the game-id range (the multi-GPU invariant)."""
    want, hist = CO.selfplay(0, n_games, seed)
    got = backend.sweep(0, n_games, seed)
    assert got[:6].tolist() == want.tolist(), (got[:6], want)
    assert got[6:].tolist() == hist.tolist()
    cuts = [0, n_games // 3, n_games // 2 + 7, n_games]
    parts = sum(backend.sweep(a, b, seed) for a, b in zip(cuts[:-1], cuts[1:]))
    assert parts.tolist() == got.tolist()
    off = backend.sweep(10_000_000_000, 10_000_000_000 + 257, seed)     # 64-bit game ids
    want_off, _ = CO.selfplay(10_000_000_000, 10_000_000_000 + 257, seed)
    assert off[:6].tolist() == want_off.tolist()


def check_step_random(backend, n_games, seed, game_base=0):
    """step_random traces == the oracle's Philox playouts; replaying the emitted (action, coin)
    trace through the oracle ends in the same state."""
    dut = backend.games(n_games)
    ref = CO.Games(n_games)
    acts, coins = [], []
    for ply in range(9):
        o = dut.step_random(seed, game_base)
        acts.append(o["action"]); coins.append(o["coin"])
        pairs = np.full((n_games, 2), -1, np.int8)
        live = o["action"] < 36
        pairs[live] = PAIRS[o["action"][live]]
        r = ref.step(pairs, o["coin"])
        assert np.array_equal(o["status"] == 2, ~live)
        assert np.array_equal(o["status"][live], r["status"][live]) and (r["status"][live] == 0).all()
        assert np.array_equal(o["reward"].view(np.uint32), r["reward"].view(np.uint32))
        assert np.array_equal(o["done"], r["done"]) and np.array_equal(o["mask"], r["mask"])
    assert_same_obs(dut.observe(), ref.observe(), "step_random final")
    fin = dut.observe()
    assert ((fin["winner"] != 0) | (fin["n_moves"] == 9)).all()        # all games terminated
    acts, coins = np.stack(acts, 1), np.stack(coins, 1)
    for g in range(min(n_games, 300)):
        w, a, c, _ = CO.playout_trace(CO.Games(1), 0, seed, game_base + g, 0)
        k = len(a)
        assert acts[g, :k].tolist() == a.tolist() and coins[g, :k].tolist() == c.tolist()
        assert (acts[g, k:] == 255).all()
        assert int(fin["winner"][g]) == w


def check_philox_coin(backend, n_games=3000, seed=99, game_base=1234):
    """step(..., choices=None): the coin is bit 0 of x1 of Philox(seed; game, len(moves), 0)."""
    rng = np.random.default_rng(seed)
    ref = CO.Games(n_games)
    dut = backend.games(n_games)
    mask = np.full(n_games, (1 << 36) - 1, np.uint64)
    for ply in range(9):
        legal = expand_mask(mask)
        k = (rng.random((n_games, 36)) * legal).argmax(1).astype(np.uint8)
        k[~legal.any(1)] = 255
        nm = ref.observe()["n_moves"]
        coins = np.array([O.policy_draw(seed, game_base + g, int(nm[g]), 0)[1] & 1
                          for g in range(n_games)], np.uint8)
        pairs = np.full((n_games, 2), -1, np.int8)
        pairs[k < 36] = PAIRS[k[k < 36]]
        r = ref.step(pairs, coins)
        o = dut.step_index(k, None, seed=seed, game_base=game_base)
        assert_same_step(o, r, f"philox coin ply {ply}")
        mask = r["mask"]
    assert_same_obs(dut.observe(), ref.observe(), "philox coin final")


def check_autoreset(backend, n_games=3000, seed=77, mode="apply", steps=40, illegal_rate=0.05):
    """qttt_step_ex with QTTT_STEP_AUTORESET / _NEXT against the oracle driven the long way: a
    game that is over on entry is replaced by a fresh game (Env.reset, env.py:55-57) and then
    stepped (apply) or left alone for this call (next).  Envs drift apart, so after a few steps
    every ply 0..8 is present in the batch at once: this is also the desynchronised-batch parity
    case (the transition's sweep is dispatched on the warp's largest len(moves))."""
    flags = {"apply": 2, "next": 4}[mode]
    rng = np.random.default_rng(seed)
    ref = CO.Games(n_games)
    fresh = CO.Games(1).raw[0].copy()
    full_mask = np.uint64((1 << 36) - 1)
    dut = backend.games(n_games)
    mask = np.full(n_games, full_mask, np.uint64)
    over = np.zeros(n_games, bool)
    resets = 0
    plies_seen = set()
    for t in range(steps):
        # actions are drawn against the position the step will see: the fresh board for envs that
        # are about to be reset
        eff_mask = np.where(over, full_mask, mask)
        legal = expand_mask(eff_mask)
        k = (rng.random((n_games, 36)) * legal).argmax(1).astype(np.uint8)
        k[~legal.any(1)] = 255
        bad = rng.random(n_games) < illegal_rate
        k[bad] = rng.integers(0, 64, int(bad.sum())).astype(np.uint8)      # any index: often illegal
        coins = rng.integers(0, 2, n_games).astype(np.uint8)
        ref.raw[over] = fresh
        plies_seen |= set(np.unique(ref.observe()["n_moves"]).tolist())
        pairs = np.full((n_games, 2), -1, np.int8)
        apply = (k < 36) & (~over if mode == "next" else np.ones(n_games, bool))
        pairs[apply] = PAIRS[k[apply]]
        r = ref.step(pairs, coins)
        o = dut.step_ex(k, coins, epoch=t + 1, flags=flags)
        where = f"{backend.name} autoreset={mode} step {t}"
        want_status = r["status"].copy()
        if mode == "next":
            want_status[over] = 0
        want_status[over] |= 4
        assert np.array_equal(o["status"], want_status), where
        assert np.array_equal(o["reward"].view(np.uint32), r["reward"].view(np.uint32)), where
        assert np.array_equal(o["done"], r["done"]) and np.array_equal(o["mask"], r["mask"]), where
        assert_same_obs(dut.observe(), ref.observe(), where)
        resets += int(over.sum())
        mask, over = r["mask"], r["done"].astype(bool)
    assert resets > n_games and plies_seen >= set(range(9))     # every env restarted; all plies mixed


def check_autoreset_random(backend, n_games=2000, seed=5, steps=30):
    """qttt_step_random_ex with QTTT_STEP_AUTORESET: continuous random self-play.  The emitted
    (action, coin) trace replayed through the oracle (with resets where the oracle says the game
    was over) reproduces every output; the draws are those of (seed, game, ply, epoch)."""
    ref = CO.Games(n_games)
    fresh = CO.Games(1).raw[0].copy()
    dut = backend.games(n_games)
    over = np.zeros(n_games, bool)
    for t in range(steps):
        ref.raw[over] = fresh
        nm = ref.observe()["n_moves"]
        lm = ref.legal_mask()
        o = dut.step_random_ex(seed, game_base=900, epoch=t + 1, flags=2)
        for g in range(0, n_games, 97):                   # spot-check the keyed draws
            x0, x1 = O.policy_draw(seed, 900 + g, int(nm[g]), ((t + 1) << 8))
            legal = [a for a in range(36) if (int(lm[g]) >> a) & 1]
            assert int(o["action"][g]) == legal[(x0 * len(legal)) >> 32] and int(o["coin"][g]) == (x1 & 1)
        pairs = PAIRS[o["action"]]
        r = ref.step(pairs, o["coin"])
        where = f"{backend.name} random autoreset step {t}"
        assert (r["status"] == 0).all(), where
        assert np.array_equal(o["status"], np.where(over, 4, 0).astype(np.uint8)), where
        assert np.array_equal(o["reward"].view(np.uint32), r["reward"].view(np.uint32)), where
        assert np.array_equal(o["done"], r["done"]) and np.array_equal(o["mask"], r["mask"]), where
        over = r["done"].astype(bool)
    assert_same_obs(dut.observe(), ref.observe(), f"{backend.name} random autoreset final")


def check_epoch_coin(backend, n_games=2000, seed=99, game_base=7):
    """The collapse coin of step(..., choices=None) is keyed (seed, game, len(moves), EPOCH):
    replaying the same actions in a later epoch draws different coins (a different Philox
    counter), each epoch matching the oracle's keyed draw.  Episode loops with reset() must not
    replay the same coins (the reference draws a fresh random.choice per collapse, qeval.py:35)."""
    rng = np.random.default_rng(seed)
    # one fixed action trace that is legal whatever the coins are: disjoint pairs then re-plays
    trace = [0, 0, 15, 15, 26, 26, 33, 33]           # (0,1)x2 (2,3)x2 (4,5)x2 (6,7)x2 -> 4 collapses
    finals = []
    for epoch in (0, 1, 2, 5):
        ref = CO.Games(n_games)
        dut = backend.games(n_games)
        for ply, a in enumerate(trace):
            k = np.full(n_games, a, np.uint8)
            nm = ref.observe()["n_moves"]
            coins = np.array([O.policy_draw(seed, game_base + g, int(nm[g]), epoch << 8)[1] & 1
                              for g in range(n_games)], np.uint8)
            r = ref.step(PAIRS[k], coins)
            if ply == 0:
                o = dut.step_ex(k, None, seed=seed, game_base=game_base, epoch=epoch, flags=1)   # reset + step
            else:
                o = dut.step_ex(k, None, seed=seed, game_base=game_base, epoch=epoch)
            assert_same_step(o, r, f"{backend.name} epoch {epoch} ply {ply}")
        obs = dut.observe()
        assert_same_obs(obs, ref.observe(), f"{backend.name} epoch {epoch} final")
        finals.append(obs["classical"].copy())
    for i in range(len(finals)):
        for j in range(i + 1, len(finals)):
            # same actions, different epoch: a large share of the games must collapse differently
            assert (finals[i] != finals[j]).any(1).mean() > 0.8
    del rng


def check_exhaustive_openings(backend, depth=3, stride=1):
    """EVERY action-index sequence of `depth` plies from the empty board (36^depth of them, legal or
    not) under every assignment of the forced coins, against the oracle after every ply: all
    placements, all 2- and 3-cycles, all illegal repeats of an opening.  `stride` thins the set."""
    idx = np.arange(0, 36 ** depth, stride, dtype=np.int64)
    acts = np.stack([(idx // 36 ** p) % 36 for p in range(depth)], 0).astype(np.uint8)      # [depth, n]
    n0 = acts.shape[1]
    for coin_bits in range(1 << depth):
        coins = np.stack([np.full(n0, (coin_bits >> p) & 1, np.uint8) for p in range(depth)], 0)
        ref = CO.Games(n0)
        dut = backend.games(n0)
        for p in range(depth):
            r = ref.step(PAIRS[acts[p]], coins[p])
            o = dut.step_index(acts[p], coins[p])
            assert_same_step(o, r, f"{backend.name} exhaustive coins={coin_bits:b} ply {p}")
        assert_same_obs(dut.observe(), ref.observe(), f"{backend.name} exhaustive coins={coin_bits:b}")
    return n0 * (1 << depth)


def check_packed_step(backend, n_games=5000, seed=31, variant=None):
    """qttt_step_packed: one byte in (action | coin << 7), one 16-bit word out (free squares |
    terminated << 9 | line << 10 | status << 11) -- same transition as the oracle, bit for bit."""
    rng = np.random.default_rng(seed)
    ref = CO.Games(n_games)
    dut = backend.games(n_games)
    mask = np.full(n_games, (1 << 36) - 1, np.uint64)
    for ply in range(11):
        legal = expand_mask(mask)
        k = (rng.random((n_games, 36)) * legal).argmax(1).astype(np.uint8)
        k[~legal.any(1)] = 63
        bad = rng.random(n_games) < 0.08
        k[bad] = rng.integers(0, 64, int(bad.sum())).astype(np.uint8)       # random index, often illegal
        coins = rng.integers(0, 2, n_games).astype(np.uint8)
        pairs = np.full((n_games, 2), -1, np.int8)
        pairs[k < 36] = PAIRS[k[k < 36]]
        r = ref.step(pairs, coins)
        if variant is None:
            res = dut.step_packed(k | (coins << 7)).astype(np.uint32)
        else:
            res = dut.step_packed(k | (coins << 7), variant=variant).astype(np.uint32)
            if variant.endswith("obs"):          # the observation that came back IS the new state
                assert np.array_equal(dut.last_obs, dut.state), f"{variant} ply {ply}: obs"
        where = f"{backend.name} packed {variant} ply {ply}"
        robs = ref.observe()
        free = ((robs["classical"] < 0) * (1 << np.arange(9))).sum(1).astype(np.uint32)
        assert np.array_equal(res & 0x1FF, free), where
        assert np.array_equal(_mask_from_obs(robs), r["mask"])               # the mask IS a function of it
        assert np.array_equal((res >> 9) & 1, r["done"].astype(np.uint32)), where
        win = (r["reward"].view(np.uint32) == 0xBF800000).astype(np.uint32)
        assert np.array_equal((res >> 10) & 1, win), where
        assert np.array_equal((res >> 11) & 3, r["status"].astype(np.uint32)), where
        assert_same_obs(dut.observe(), robs, where)
        mask = r["mask"]


def check_golden_features(backend):
    """GameState.to_vector (mcts.py:67-85) recorded from the live reference; float32 vs the
    reference's float64 within 1e-6."""
    recs = load_golden("features_v1.json.gz")
    n = len(recs)
    classical = np.array([r["board"] for r in recs], np.int8)
    moves = np.full((n, 9, 2), -1, np.int8)
    n_moves = np.array([len(r["moves"]) for r in recs], np.uint8)
    for i, r in enumerate(recs):
        for m in r["moves"]:
            moves[i, m[2]] = m[:2]
    got = backend.games(n).load(classical, moves, n_moves).features().reshape(n, 180).astype(np.float64)
    want = np.zeros((n, 180))
    for i, r in enumerate(recs):
        for idx, v in r["nonzero"]:
            want[i, idx] = v
    assert np.abs(got - want).max() <= 1e-6
    assert ((got != 0) == (want != 0)).all()


def check_golden_getmask(backend):
    """nn.Model.get_mask (nn.py:44-61) recorded from the live reference (oracle/make_golden_getmask.py),
    next to the reference's own GameState.action_mask() of the same nodes."""
    recs = load_golden("getmask_v1.json.gz")
    n = len(recs)
    classical = np.array([r["board"] for r in recs], np.int8)
    moves = np.full((n, 9, 2), -1, np.int8)
    n_moves = np.array([len(r["moves"]) for r in recs], np.uint8)
    for i, r in enumerate(recs):
        for m in r["moves"]:
            moves[i, m[2]] = m[:2]
    games = backend.games(n).load(classical, moves, n_moves)
    got = games.get_mask()
    want = np.array([r["get_mask"] for r in recs], bool)
    assert got.shape == (n, 36) and np.array_equal(got, want)
    legal = games.observe()["mask_bool"].astype(bool)
    assert np.array_equal(legal, np.array([r["action_mask"] for r in recs], bool))


def check_step_features(backend, n_games=3000, seed=41):
    """qttt_step_features == qttt_step followed by to_vector / get_mask of the new state, on a
    desynchronised (auto-resetting) batch with some illegal actions; the step outputs against the
    oracle, the encodings against the stand-alone encoders of the same backend (which the golden
    tests pin to the reference)."""
    rng = np.random.default_rng(seed)
    ref = CO.Games(n_games)
    fresh = CO.Games(1).raw[0].copy()
    full_mask = np.uint64((1 << 36) - 1)
    dut = backend.games(n_games)
    mask = np.full(n_games, full_mask, np.uint64)
    over = np.zeros(n_games, bool)
    for t in range(24):
        legal = expand_mask(np.where(over, full_mask, mask))
        k = (rng.random((n_games, 36)) * legal).argmax(1).astype(np.uint8)
        bad = rng.random(n_games) < 0.05
        k[bad] = rng.integers(0, 64, int(bad.sum())).astype(np.uint8)
        coins = rng.integers(0, 2, n_games).astype(np.uint8)
        ref.raw[over] = fresh
        pairs = np.full((n_games, 2), -1, np.int8)
        pairs[k < 36] = PAIRS[k[k < 36]]
        r = ref.step(pairs, coins)
        o = dut.step_features(k, coins, epoch=t + 1, flags=2)
        where = f"{backend.name} step_features step {t}"
        assert np.array_equal(o["reward"].view(np.uint32), r["reward"].view(np.uint32)), where
        assert np.array_equal(o["done"], r["done"]) and np.array_equal(o["mask"], r["mask"]), where
        assert_same_obs(dut.observe(), ref.observe(), where)
        assert np.array_equal(o["features"], dut.features()), where
        assert np.array_equal(o["illegal_mask"], dut.get_mask()), where
        assert np.array_equal(o["illegal_mask"], ~expand_mask(r["mask"]).astype(bool)), where
        mask, over = r["mask"], r["done"].astype(bool)


def check_step_obs(backend, n_games=3000, seed=43):
    """qttt_step_obs: the step outputs and the observation that comes out of the same launch,
    against the oracle after every step -- plain mode with (a, b) pairs played to the end and
    beyond (post-terminal moves, illegal actions), then an auto-resetting batch in both modes
    (every ply present in the batch at once), some illegal actions throughout."""
    rng = np.random.default_rng(seed)
    full_mask = np.uint64((1 << 36) - 1)
    fresh = CO.Games(1).raw[0].copy()

    def same(o, r, ref, where):
        assert np.array_equal(o["reward"].view(np.uint32), r["reward"].view(np.uint32)), where
        assert np.array_equal(o["done"], r["done"]) and np.array_equal(o["mask"], r["mask"]), where
        want = ref.observe()
        for k in ("classical", "q_p1", "q_p2", "turn"):
            assert np.array_equal(o["obs"][k], want[k]), f"{where}: {k}"

    ref, dut = CO.Games(n_games), backend.games(n_games)
    mask = np.full(n_games, full_mask, np.uint64)
    for t in range(11):
        legal = expand_mask(mask)
        k = (rng.random((n_games, 36)) * legal).argmax(1)
        pairs = PAIRS[k].copy()
        flip = rng.random(n_games) < 0.5
        pairs[flip] = pairs[flip][:, ::-1]
        bad = rng.random(n_games) < 0.05
        pairs[bad] = rng.integers(-1, 10, (int(bad.sum()), 2)).astype(np.int8)
        coins = rng.integers(0, 2, n_games).astype(np.uint8)
        r = ref.step(pairs, coins)
        o = dut.step_obs(pairs, coins, pairs=True)
        same(o, r, ref, f"{backend.name} step_obs pairs step {t}")
        assert_same_obs(dut.observe(), ref.observe(), f"{backend.name} step_obs pairs state {t}")
        mask = r["mask"]
    for flags in (2, 4):
        ref, dut = CO.Games(n_games), backend.games(n_games)
        mask = np.full(n_games, full_mask, np.uint64)
        over = np.zeros(n_games, bool)
        for t in range(20):
            legal = expand_mask(np.where(over, full_mask, mask))
            k = (rng.random((n_games, 36)) * legal).argmax(1).astype(np.uint8)
            bad = rng.random(n_games) < 0.05
            k[bad] = rng.integers(0, 64, int(bad.sum())).astype(np.uint8)
            coins = rng.integers(0, 2, n_games).astype(np.uint8)
            ref.raw[over] = fresh
            pairs = np.full((n_games, 2), -1, np.int8)
            play = (k < 36) & ~(over if flags == 4 else np.zeros(n_games, bool))   # "next": a reset env ignores its action
            pairs[play] = PAIRS[k[play]]
            r = ref.step(pairs, coins)
            o = dut.step_obs(k, coins, epoch=t + 1, flags=flags)
            same(o, r, ref, f"{backend.name} step_obs flags {flags} step {t}")
            mask, over = r["mask"], r["done"].astype(bool)


def _pack_oracle_games(backend, games):
    """oracle Game objects -> packed states via the backend's pack."""
    n = len(games)
    classical = np.array([g.board for g in games], np.int8)
    moves = np.full((n, 9, 2), -1, np.int8)
    nm = np.array([len(g.moves) for g in games], np.uint8)
    for i, g in enumerate(games):
        for a, b, idx in g.moves:
            moves[i, idx] = (a, b)
    return np.asarray(backend.games(n).load(classical, moves, nm).state).copy()


def check_golden_mcts_search(backend):
    """Search statistics recorded from the live reference MCTS (keyed stream): N, Q (float64,
    bit-exact), Ntot and choose() after every contemplate, across sync()s."""
    cases = load_golden("mcts_search_v1.json.gz")
    for case in cases:
        g = O.Game()
        for a, b, c in case["prefix"]:
            g.place(a, b, lambda: c)
        budget = sum(st["rollouts"] for st in case["stages"])
        s = backend.mcts(_pack_oracle_games(backend, [g]), case["num_simulations"], case["seed"],
                         case["root_index"], budget)
        created, syncs = 1, 0
        for st in case["stages"]:
            before = int(s.live()[0])
            s.contemplate(st["rollouts"])
            created += int(s.live()[0]) - before
            n, q, ntot, ch = s.stats()
            assert n[0].tolist() == st["N"], case["root_index"]
            assert q[0].tolist() == st["Q"], case["root_index"]          # float64, exact
            assert int(ntot[0]) == st["Ntot"] and int(ch[0]) == st["choose"]
            if st["move"] is None:
                break
            act, c = st["move"]
            a, b = O.PAIRS[act]
            g.place(a, b, lambda: c)
            before = int(s.live()[0])
            s.sync(np.array([act], np.uint8), _pack_oracle_games(backend, [g]))
            syncs += 1
            # MCTS._prune (mcts.py:222-231, 330-337): the old root and the siblings' subtrees are gone
            assert int(s.live()[0]) < before, case["root_index"]
        assert int(s.errors()[0]) == 0
        # nodes reclaimed by sync are reused: the pool's high-water mark stays below the number of
        # nodes ever created (and the statistics above stayed bit-identical all the same)
        if syncs >= 2:
            assert int(s.taken()[0]) < created, (case["root_index"], int(s.taken()[0]), created)


def check_mcts_batch_vs_oracle(backend, n_roots=24, rollouts=80, sims=8, seed=4242, root_base=100):
    """A batch of different mid-game roots searched concurrently == the oracle searching each
    root alone (root_base + index keys the stream)."""
    from oracle import mcts_oracle as MO
    import random
    rng = random.Random(seed)
    games = []
    while len(games) < n_roots:
        g = O.Game()
        for _ in range(rng.randrange(0, 6)):
            if g.terminal():
                break
            a, b = O.PAIRS[rng.choice(g.legal_actions())]
            c = rng.randrange(2)
            g.place(a, b, lambda: c)
        games.append(g)                       # terminal roots included on purpose
    s = backend.mcts(_pack_oracle_games(backend, games), sims, seed, root_base, rollouts)
    s.contemplate(rollouts // 2)
    s.contemplate(rollouts - rollouts // 2)   # two calls continue the same stream
    n, q, ntot, ch = s.stats()
    for i, g in enumerate(games):
        m = MO.MCTS(rollouts, sims, seed, root_base + i)
        m.reset(g)
        m.contemplate()
        wn, wq, wt = m.root_stats()
        assert n[i].tolist() == wn and q[i].tolist() == wq and int(ntot[i]) == wt, i
        want = m.choose()
        assert int(ch[i]) == (255 if want is None else want)
    assert (s.errors() == 0).all()


def check_rollout_frequencies_vs_reference_simulate(backend, n_rollouts=4096, seed=5):
    """config 4, statistical bridge: per-root X / O / draw frequencies of the Philox-driven
    rollouts vs counts recorded from the UNMODIFIED reference's MCTS._simulate with its own
    randomness (tests/golden/simulate_freq_v1.json), 4.5 sigma per cell."""
    cases = load_golden("simulate_freq_v1.json")
    games = []
    for case in cases:
        g = O.Game()
        for a, b, c in case["prefix"]:
            g.place(a, b, lambda: c)
        games.append(g)
    n = len(games)
    classical = np.array([g.board for g in games], np.int8)
    moves = np.full((n, 9, 2), -1, np.int8)
    nm = np.array([len(g.moves) for g in games], np.uint8)
    for i, g in enumerate(games):
        for a, b, idx in g.moves:
            moves[i, idx] = (a, b)
    tallies, _, _ = backend.games(n).load(classical, moves, nm).rollout(n_rollouts, seed)
    for i, case in enumerate(cases):
        m = case["playouts"]
        for k, key in enumerate(("x", "o", "draw")):
            p, q = tallies[i, k] / n_rollouts, case[key] / m
            pooled = (tallies[i, k] + case[key]) / (n_rollouts + m)
            sigma = max((pooled * (1 - pooled) * (1 / n_rollouts + 1 / m)) ** 0.5, 1e-9)
            assert abs(p - q) < 4.5 * sigma or abs(p - q) < 1e-12, (i, key, p, q)
