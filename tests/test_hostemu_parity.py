"""CPU (no GPU): the CUDA source's per-game functions, compiled for the host
(tests/hostemu), against the oracle and the reference fixtures.  This is a pre-flight of the
kernel LOGIC; the real parity tests are tests/test_gpu_parity.py, which run the same suite
through the C ABI on the device."""
import pytest

import parity_suite as S
from backends import EmuBackend


@pytest.fixture(scope="module")
def emu():
    return EmuBackend()


def test_golden_traces(emu):
    S.check_golden_traces(emu)


def test_golden_qeval(emu):
    S.check_golden_qeval(emu)


def test_golden_mcts_step(emu):
    S.check_golden_mcts_step(emu)


@pytest.mark.parametrize("seed,illegal,overrun,fmt", [
    (1, 0.0, False, "pair"), (2, 0.1, False, "pair"), (3, 0.1, True, "pair"),
    (4, 0.0, False, "index"), (5, 0.1, True, "index")])
def test_random_play(emu, seed, illegal, overrun, fmt):
    steps = S.check_random_play(emu, 20000, seed, illegal, overrun, fmt)
    assert steps > 100000


def test_pack_observe_roundtrip(emu):
    S.check_pack_observe_roundtrip(emu)


def test_qeval_both(emu):
    S.check_qeval_both(emu, 20000, 7)


def test_rollout(emu):
    S.check_rollout(emu, 64, 32, 11)
    S.check_rollout_terminal_roots(emu)


def test_sweep(emu):
    S.check_sweep(emu, 30000, 13)


def test_step_random(emu):
    S.check_step_random(emu, 2000, 17, game_base=5)


def test_philox_coin(emu):
    S.check_philox_coin(emu, 1500)


def test_packed_step(emu):
    S.check_packed_step(emu)


def test_golden_features(emu):
    S.check_golden_features(emu)


def test_render_text_matches_reference_display():
    import qtttgym_b200 as Q
    from helpers import load_golden
    for r in load_golden("features_v1.json.gz"):
        moves = [[-1, -1]] * 9
        for m in r["moves"]:
            moves[m[2]] = m[:2]
        # displayBoard prints the string followed by print()'s own newline
        assert Q.render_text(r["board"], moves, len(r["moves"])) + "\n" == r["display"]


def test_golden_mcts_search(emu):
    S.check_golden_mcts_search(emu)


def test_mcts_batch_vs_oracle(emu):
    S.check_mcts_batch_vs_oracle(emu)


def test_rollout_frequencies_vs_reference_simulate(emu):
    S.check_rollout_frequencies_vs_reference_simulate(emu, n_rollouts=2048)


@pytest.mark.parametrize("mode", ["apply", "next"])
def test_autoreset(emu, mode):
    S.check_autoreset(emu, 1500, 77, mode)


def test_autoreset_random(emu):
    S.check_autoreset_random(emu, 1000, 5)


def test_epoch_coin(emu):
    S.check_epoch_coin(emu, 600)


def test_golden_getmask(emu):
    S.check_golden_getmask(emu)


def test_step_features(emu):
    S.check_step_features(emu, 1000)


def test_step_obs(emu):
    S.check_step_obs(emu, 700)


def test_exhaustive_openings(emu):
    assert S.check_exhaustive_openings(emu, depth=3, stride=7) > 50000
