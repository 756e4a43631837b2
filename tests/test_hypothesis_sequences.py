"""Property test (CPU): ARBITRARY action sequences -- legal, illegal, repeated, off-board,
negative, after termination -- through the CUDA source's transition (host emulation) and the
Python oracle must agree on every observable after every step (quirks Q1-Q3, Q7)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import qttt_oracle as O
from oracle.tracegen import f32_bits

from backends import EmuBackend

square = st.one_of(st.integers(0, 8), st.integers(-3, 12), st.integers(-128, 127))
step = st.tuples(square, square, st.integers(0, 1))


@settings(max_examples=300, deadline=None)
@given(st.lists(step, min_size=1, max_size=16))
def test_pair_sequences_match_oracle(seq):
    emu = EmuBackend().games(1)
    env = O.Env()
    env.reset()
    for a, b, coin in seq:
        obs, r, term, _, _ = env.step((a, b), coin=lambda: coin)
        out = emu.step(np.array([[a, b]], np.int8), np.array([coin], np.uint8))
        got = emu.observe()
        g = env.game
        assert int(out["reward"].view(np.uint32)[0]) == f32_bits(r)
        assert bool(out["done"][0]) == term
        assert int(out["mask"][0]) == g.legal_mask()
        assert int(out["status"][0]) == env.last_status
        assert got["classical"][0].tolist() == g.board
        nm = int(got["n_moves"][0])
        assert nm == len(g.moves)
        assert [tuple(m) + (i,) for i, m in enumerate(got["moves"][0][:nm].tolist())] == g.moves
        assert got["rounds"][0].tolist() == list(g.win_rounds())
        assert int(got["winner"][0]) == g.winner() and float(got["reward_p1"][0]) == g.reward_p1()
        assert [tuple(p) for p in got["q_p1"][0].tolist() if p[0] >= 0] == obs["q_states_p1"]
        assert [tuple(p) for p in got["q_p2"][0].tolist() if p[0] >= 0] == obs["q_states_p2"]


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 255), st.integers(0, 1)), min_size=1, max_size=16))
def test_index_sequences_match_oracle(seq):
    emu = EmuBackend().games(1)
    env = O.Env()
    env.reset()
    for idx, coin in seq:
        pair = O.PAIRS[idx] if idx < 36 else (-1, -1)
        _, r, term, _, _ = env.step(pair, coin=lambda: coin)
        out = emu.step_index(np.array([idx], np.uint8), np.array([coin], np.uint8))
        assert int(out["reward"].view(np.uint32)[0]) == f32_bits(r) and bool(out["done"][0]) == term
        assert int(out["mask"][0]) == env.game.legal_mask() and int(out["status"][0]) == env.last_status
        assert emu.observe()["classical"][0].tolist() == env.game.board
