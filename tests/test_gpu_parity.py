"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes ->
libqttt_b200.so), against the oracle and the fixtures recorded from the live reference."""
import os

import numpy as np
import pytest

import parity_suite as S
from backends import CudaBackend
from helpers import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch
    assert torch.cuda.is_available()
    import qtttgym_b200._lib as L
    L.lib()                       # fails loudly when the CUDA library is missing
    return CudaBackend()


def test_native_library_is_the_in_tree_one(cuda):
    import qtttgym_b200._lib as L
    with open("/proc/self/maps") as f:
        loaded = [l.split()[-1] for l in f if "libqttt_b200.so" in l]
    assert loaded and os.path.samefile(loaded[0], L.LIB)
    assert not any("libqttt_oracle" in p and "qtttgym_b200" in p for p in loaded)


def test_golden_traces(cuda):
    S.check_golden_traces(cuda)


def test_golden_qeval(cuda):
    S.check_golden_qeval(cuda)


def test_golden_mcts_step(cuda):
    S.check_golden_mcts_step(cuda)


@pytest.mark.parametrize("seed,illegal,overrun,fmt", [
    (20261018, 0.0, False, "index"),      # config 2: 4096 envs, random legal actions, forced coins
    (20261019, 0.1, True, "pair"),        # config 2b: 10 % illegal actions + post-terminal moves
    (20261020, 0.1, True, "index")])
def test_config2_4096_envs(cuda, seed, illegal, overrun, fmt):
    S.check_random_play(cuda, 4096, seed, illegal, overrun, fmt)


def test_equivalence_suite_one_million_games(cuda):
    """North-star: bit-exact on 1e6 random games (4 x 262,144 envs, every output after every ply)."""
    total = 0
    for b in range(4):
        total += S.check_random_play(cuda, 262144, 1000 + b, illegal_rate=0.02 * (b % 2), fmt=("pair", "index")[b // 2])
    assert total > 8.0e6


@pytest.mark.parametrize("n", [1, 31, 257, 4097])
def test_ragged_batch_sizes(cuda, n):
    S.check_random_play(cuda, n, 77 + n, illegal_rate=0.05, overrun=True)


def test_empty_batch(cuda):
    import torch
    import qtttgym_b200 as Q
    env = Q.BatchedEnv(0)
    obs, r, term, trunc, info = env.step(torch.zeros(0, dtype=torch.uint8, device="cuda"))
    assert r.numel() == 0 and term.numel() == 0
    assert Q.selfplay_sweep(5, 5, 1).sum().item() == 0


def test_pack_observe_roundtrip(cuda):
    S.check_pack_observe_roundtrip(cuda, 20000)


def test_config3_qeval_one_million_boards(cuda):
    S.check_qeval_both(cuda, 1 << 20, 7)


def test_config4_rollout_1024_roots_x_256(cuda):
    S.check_rollout(cuda, 1024, 256, 11)
    S.check_rollout_terminal_roots(cuda)


def test_rollout_non_multiple_of_block(cuda):
    S.check_rollout(cuda, 33, 300, 12)
    S.check_rollout(cuda, 5, 1, 13)


def test_config5_sweep_matches_oracle_and_shards(cuda):
    S.check_sweep(cuda, 2_000_000, 13)


def test_step_random_traces(cuda):
    S.check_step_random(cuda, 50000, 17, game_base=123456789012)


def test_philox_coin(cuda):
    S.check_philox_coin(cuda, 3000)


def test_step_api_and_sweep_agree_at_full_size(cuda):
    """Size-independent property at 2^24 envs: playing the random policy through the step API
    (9 launches of K1) gives exactly the tallies of the fused sweep (K5) on the same game ids."""
    import torch
    import qtttgym_b200 as Q
    n, seed, base = 1 << 24, 424242, 7_000_000_000
    env = Q.BatchedEnv(n, seed=seed, game_base=base)
    steps = 0
    for _ in range(9):
        _, _, _, _, info = env.step_random()
        steps += int((info["status"] == 0).sum().item())
    obs = env.observation(extras=True)
    w = torch.bincount(obs["winner"].long(), minlength=3)
    stats = Q.selfplay_sweep(base, base + n, seed)
    assert [int(w[1]), int(w[2]), int(w[0])] == stats[:3].tolist()
    assert steps == int(stats[3]) and int(stats[5]) == n
    assert bool(((obs["winner"] != 0) | (obs["n_moves"] == 9)).all())


def test_population_statistics(cuda):
    """T2: tallies of 2M Philox games vs the reference's own MT19937 tallies (20k games)."""
    ref = load_golden("population_v1.json")
    stats = cuda.sweep(0, 2_000_000, 5)
    n, m = stats[5], ref["games"]
    for k, key in enumerate(("x", "o", "draw")):
        p, q = stats[k] / n, ref[key] / m
        assert abs(p - q) < 5 * (q * (1 - q) * (1 / n + 1 / m)) ** 0.5
    hist = stats[6:] / n
    ref_hist = np.array(ref["steps_hist"]) / m
    assert np.abs(hist - ref_hist).max() < 0.02


def test_single_env_adapter_reference_types(cuda):
    """Env (num_envs=1) returns the reference's Python types and values on the Appendix-A games."""
    import qtttgym_b200 as Q
    kat = load_golden("kat_appendix_a.json.gz")
    for name, g in kat.items():
        env = Q.Env()
        obs, info = env.reset(seed=123)
        assert info == {} and obs == {"q_states_p1": [], "q_states_p2": [], "classical": [-1] * 9, "turn": 0}
        for (a, b, coin), rec in zip(g["trace"], g["records"]):
            obs, r, term, trunc, info = env.step((a, b), coin=coin)
            assert isinstance(r, float) and isinstance(term, bool) and trunc is False and info == {}
            assert np.float32(r).view(np.uint32) == rec["reward_bits"], name
            assert term == rec["terminated"]
            assert obs["classical"] == rec["board"] and obs["turn"] == rec["turn"]
            assert [list(p) for p in obs["q_states_p1"]] == rec["q1"]
            assert [list(p) for p in obs["q_states_p2"]] == rec["q2"]
            assert env.turn() == len(rec["moves"]) and env._reward() == rec["reward_p1"]


def test_reference_user_code_runs_unchanged(cuda):
    """A loop written against qtttgym.Env -- including env.action_space / observation_space --
    runs unchanged on the drop-in adapter, and two Env instances do not share collapse coins."""
    import qtttgym_b200 as Q

    def user_code(env, episodes):
        total, boards = 0, []
        for _ in range(episodes):
            obs, info = env.reset()
            assert env.observation_space.contains({k: (np.asarray(v, np.int32) if k == "classical" else v)
                                                   for k, v in obs.items()})
            terminated = False
            while not terminated:
                action = env.action_space.sample()
                assert env.action_space.contains(action) and action in env.action_space
                obs, r, terminated, truncated, info = env.step(action)
                assert isinstance(obs["classical"], list) and isinstance(r, float) and not truncated
                total += 1
            boards.append(tuple(obs["classical"]))
        return total, boards

    env = Q.Env()
    steps, _ = user_code(env, 20)
    assert steps >= 100
    assert len(env.action_space) == 2 and env.action_space[0].n == 9
    # same actions, consecutive episodes and independent instances: the collapses differ
    outcomes = set()
    for inst in range(6):
        e = Q.Env()
        for _ in range(2):
            e.reset()
            for a in [(0, 1), (1, 0), (2, 3), (3, 2), (4, 5), (5, 4)]:
                obs, *_ = e.step(a)
            outcomes.add(tuple(obs["classical"]))
    assert len(outcomes) > 3
    # a fixed seed reproduces the first episode
    a, b = Q.Env(seed=5), Q.Env(seed=5)
    for mv in [(0, 1), (1, 0), (2, 3), (3, 2)]:
        assert a.step(mv)[0] == b.step(mv)[0]


def test_vector_env_next_step_autoreset(cuda):
    """gymnasium.vector conventions: spaces, (obs, info) / 5-tuple shapes, next-step autoreset."""
    import torch
    import qtttgym_b200 as Q
    n = 4096
    venv = Q.VectorEnv(n, seed=3, reward="p1")
    assert venv.num_envs == n and venv.single_action_space.contains((0, 8))
    obs, info = venv.reset(seed=1)
    assert obs["classical"].shape == (n, 9) and obs["q_states_p1"].shape == (n, 5, 2) and obs["turn"].shape == (n,)
    assert bool((obs["classical"] == -1).all())
    episodes = torch.zeros(n, dtype=torch.int64, device="cuda")
    prev_term = torch.zeros(n, dtype=torch.bool, device="cuda")
    returns = torch.zeros(3, dtype=torch.int64, device="cuda")
    for t in range(60):
        actions = venv.sample_actions()
        obs, reward, terminated, truncated, info = venv.step(actions)
        # envs that terminated at the previous step were reset by this call: fresh board, no reward
        assert bool(((obs["classical"][prev_term] == -1).all(1)).all())
        assert bool((obs["turn"][prev_term] == 0).all()) and not bool(terminated[prev_term].any())
        assert bool((reward[prev_term] == 0).all()) and bool(info["reset"][prev_term].all())
        assert not bool(info["reset"][~prev_term].any()) and not bool(truncated.any())
        assert not bool(info["invalid"][~prev_term].any())              # sampled actions are legal
        episodes += terminated
        returns += torch.bincount((reward[terminated] + 1).long(), minlength=3)
        prev_term = terminated.clone()
    assert int(episodes.min()) >= 4                                     # every env played several episodes
    tot = int(returns.sum())
    x, o, d = int(returns[2]) / tot, int(returns[0]) / tot, int(returns[1]) / tot
    assert abs(x - 0.5816) < 0.02 and abs(o - 0.2914) < 0.02 and abs(d - 0.1271) < 0.02   # SURVEY 8(d) population


def test_qeval_plugin_seam(cuda):
    """QEvalB200.eval is a drop-in for QEvalClassic.eval (board.py:51 -> qeval.py:5)."""
    import qtttgym_b200 as Q
    ev = Q.QEvalB200()
    for case in load_golden("qeval_v1.json.gz")[:60]:
        ent = [tuple(m) for m in case["entangled"]]
        ev.force(0)
        assert ev.eval(ent) == case["out0"]
        ev.force(1)
        assert ev.eval(ent) == case["out1"]
    # unforced: the coin comes from the stdlib random module, like the reference
    import random
    random.seed(3)
    got = {tuple(ev.eval([(0, 1, 0), (0, 1, 1)])) for _ in range(40)}
    assert got == {(0, 1), (1, 0)}


def test_square_probabilities(cuda):
    import torch
    import qtttgym_b200 as Q
    cl, mv, nm, act = S.harvest_positions(5000, 3, "closing")
    st = Q.pack_states(cl, mv, nm)
    res = Q.qeval_both(st, torch.from_numpy(act).cuda(), want_squares=True)
    p = Q.square_probabilities(res["sq0"], res["sq1"]).cpu().numpy()
    assert set(np.unique(p).tolist()) <= {0.0, 0.5, 1.0}
    rows = p.sum(2)                                # each measured move: total probability 1
    measured = (res["sq0"] >= 0).cpu().numpy()
    assert np.abs(rows[measured] - 1.0).max() <= 1e-6 and np.abs(rows[~measured]).max() == 0
    cols = p.sum(1)                                # each square of the component gets exactly one move
    assert set(np.unique(cols).tolist()) <= {0.0, 1.0}


def test_error_codes(cuda):
    import torch
    import qtttgym_b200._lib as L
    lib = L.lib()
    st = torch.zeros((8, 4), dtype=torch.int32, device="cuda")
    act = torch.zeros(8, dtype=torch.uint8, device="cuda")
    assert lib.qttt_reset(None, None, 8, None) == -1
    assert lib.qttt_reset(st.data_ptr(), None, -1, None) == -1
    assert lib.qttt_step(st.data_ptr(), act.data_ptr(), 7, None, 0, 0, None, None, None, None, 8, None) == -1
    assert lib.qttt_reset(st.data_ptr() + 4, None, 4, None) == -2
    assert lib.qttt_rollout(st.data_ptr(), 8, 0, 0, None, None, None, None) == -1
    assert lib.qttt_sweep(5, 4, 0, st.data_ptr(), None) == -1
    assert b"invalid argument" in lib.qttt_strerror(-1) and b"aligned" in lib.qttt_strerror(-2)
    with pytest.raises(RuntimeError):
        L.check(-1)
    torch.cuda.synchronize()


def test_observe_output_subsets_and_alignment(cuda):
    """qttt_observe dispatches on the set of outputs and on their alignment: every path gives
    the same bytes (mid-game states incl. collapses, a size that is not a multiple of 256)."""
    import torch
    import qtttgym_b200 as Q
    import qtttgym_b200._lib as L
    n = 3 * 256 + 77
    env = Q.BatchedEnv(n, seed=17)
    for _ in range(6):
        env.step_random()
    full = Q.observe_states(env.state, extras=True)                # every output, aligned
    lean = Q.observe_states(env.state)                             # the env.py set
    for k in lean:
        assert torch.equal(lean[k], full[k]), k
    # outputs that start one row late: odd byte offsets, the byte-granular copy-out
    off = {k: torch.zeros((n + 1,) + tuple(v.shape[1:]), dtype=torch.uint8 if v.dtype == torch.bool else v.dtype,
                          device=v.device)[1:] for k, v in full.items()}
    got = Q.observe_states(env.state, extras=True, out=off)
    for k, v in full.items():
        assert torch.equal(got[k].to(v.dtype), v), k
    off2 = {k: off[k] for k in lean}
    got2 = Q.observe_states(env.state, out=off2)
    for k in lean:
        assert torch.equal(got2[k], full[k]), k
    # single outputs (the generic specialisation)
    assert torch.equal(env.turn(), full["n_moves"])
    assert torch.equal(env.action_mask(), full["action_mask"])
    assert torch.equal(env.reward_p1(), full["reward_p1"])
    assert torch.equal(env.winner(), full["winner"])
    # q_states_p2 and rounds are written with 8- / 2-byte stores
    lib = L.lib()
    buf = torch.zeros(16 * n + 64, dtype=torch.uint8, device="cuda")
    none = [None] * 4
    assert lib.qttt_observe(env.state.data_ptr(), None, None, None, None, buf.data_ptr() + 4, *none, None, n, None) == -2
    assert lib.qttt_observe(env.state.data_ptr(), None, None, None, None, None, None, buf.data_ptr() + 1,
                            None, None, None, n, None) == -2
    torch.cuda.synchronize()


def test_step_host_equals_step(cuda):
    """The end-to-end entry (pinned host buffers, chunk-pipelined) gives the same results."""
    import torch
    import qtttgym_b200 as Q
    n = 100_003
    gen = Q.BatchedEnv(n, seed=9)
    a = Q.BatchedEnv(n, seed=9)
    b = Q.BatchedEnv(n, seed=9)
    hr = torch.empty(n, dtype=torch.float32).pin_memory()
    hd = torch.empty(n, dtype=torch.bool).pin_memory()
    hm = torch.empty(n, dtype=torch.int64).pin_memory()
    for ply in range(9):
        info = gen.step_random(record=True)[4]
        act, coin = info["action"].clone(), info["coin"].clone()
        _, r, t, _, i2 = a.step(act, coin)
        b.step_host(act.cpu().pin_memory(), coin.cpu().pin_memory(), hr, hd, hm, chunks=5, n_streams=3)
        torch.cuda.synchronize()
        assert torch.equal(r.cpu().view(torch.int32), hr.view(torch.int32))
        assert torch.equal(t.cpu(), hd) and torch.equal(i2["action_mask"].cpu(), hm)
        assert torch.equal(a.state, b.state) and torch.equal(a.state, gen.state)


def test_exhaustive_openings(cuda):
    """all 36^3 three-ply openings x all 8 coin assignments (373,248 games) and a thinned set of the
    36^4 four-ply ones x 16 coin assignments, every output after every ply"""
    assert S.check_exhaustive_openings(cuda, depth=3) == 36 ** 3 * 8
    assert S.check_exhaustive_openings(cuda, depth=4, stride=5) > 5_000_000


def test_packed_step_host(cuda):
    S.check_packed_step(cuda, 40_001)


@pytest.mark.parametrize("n", [1, 6, 255, 20_003, 70_000])
def test_packed12_step_host(cuda, n):
    """bit-packed results (12 bits per env, four envs in three words): sizes that are not
    multiples of 4 / 256 / the block span, and one that is"""
    S.check_packed_step(cuda, n, variant="mapped12")
    S.check_packed_step(cuda, n, variant="copy12")


@pytest.mark.parametrize("n", [1, 777, 40_003])
def test_step_host_obs12_equals_step_and_observation(cuda, n):
    """The 12-byte observation records (qttt_step_packed_host_obs12 + unpack_obs12) carry exactly what
    step() and observation() return: games played to the end and beyond, some illegal actions."""
    import torch
    import qtttgym_b200 as Q
    a, b = Q.BatchedEnv(n, seed=4), Q.BatchedEnv(n, seed=4)
    a.reset(), b.reset()
    rec = torch.empty((n, 3), dtype=torch.int32).pin_memory()
    g = torch.Generator(device="cuda").manual_seed(3)
    for ply in range(11):
        legal = b.action_mask().float()
        legal[legal.sum(1) == 0, 0] = 1.0
        act = torch.multinomial(legal, 1, generator=g).squeeze(1).to(torch.uint8)
        bad = torch.rand(n, device="cuda", generator=g) < 0.06
        act = torch.where(bad, torch.randint(0, 36, (n,), device="cuda", generator=g).to(torch.uint8), act)
        coin = torch.randint(0, 2, (n,), device="cuda", generator=g).to(torch.uint8)
        _, reward, term, _, info = b.step(act, coin)
        want = b.observation()
        a.step_host_obs12(Q.pack_actions(act, coin).cpu().pin_memory(), rec, chunks=3, n_streams=2)
        torch.cuda.synchronize()
        obs, r2, t2, m2, s2 = Q.unpack_obs12(rec)
        assert torch.equal(a.state, b.state), ply
        for k in want:
            assert torch.equal(obs[k], want[k].cpu()), (ply, k)
        assert torch.equal(r2.view(torch.int32), reward.cpu().view(torch.int32)) and torch.equal(t2, term.cpu()), ply
        assert torch.equal(m2, info["action_mask"].cpu()) and torch.equal(s2, (info["status"] & 1).cpu()), ply


@pytest.mark.parametrize("variant", ["copy_obs", "mapped", "mapped_obs"])
def test_packed_step_host_variants(cuda, variant):
    """the observation coming back with the result words, and the zero-copy path (the kernel
    reads / writes the pinned host buffers itself) -- ragged size: full and partial 256-chunks"""
    S.check_packed_step(cuda, 20_003, variant=variant)


@pytest.mark.parametrize("mode", ["apply", "next"])
def test_autoreset_desynchronised_batch(cuda, mode):
    S.check_autoreset(cuda, 30_000, 77, mode)


def test_autoreset_random_selfplay(cuda):
    S.check_autoreset_random(cuda, 20_000, 5)


def test_epoch_separates_episodes(cuda):
    S.check_epoch_coin(cuda, 3000)


def test_reset_between_episodes_changes_the_coins(cuda):
    """an episode loop with env.reset() must not replay the same Philox coins."""
    import torch
    import qtttgym_b200 as Q
    n = 4096
    env = Q.BatchedEnv(n, seed=0)
    trace = [0, 0, 15, 15, 26, 26, 33, 33]
    boards = []
    for episode in range(3):
        if episode:
            env.reset()
        assert env.epoch == episode
        assert not bool(env.done.any()) and not bool(env.status.any())       # reset clears the step outputs
        assert bool((env.reward.view(torch.int32) == -2147483648).all())     # -0.0f
        for a in trace:
            env.step(torch.full((n,), a, dtype=torch.uint8, device="cuda"))
        boards.append(env.observation()["classical"].clone())
    assert float((boards[0] != boards[1]).any(1).float().mean()) > 0.8
    assert float((boards[1] != boards[2]).any(1).float().mean()) > 0.8
    sd = env.state_dict()
    assert sd["epoch"] == 2


def test_info_keys_and_out_of_range_actions(cuda):
    import torch
    import qtttgym_b200 as Q
    env = Q.BatchedEnv(5, seed=1)
    # int64 indices outside 0..35 must be illegal, not wrap modulo 256 (260 -> 4)
    _, _, _, _, info = env.step(torch.tensor([260, -1, 36, 4, 0], device="cuda"))
    assert info["invalid"].tolist() == [True, True, True, False, False]
    assert "winner" in info and "reward_p1" in info and "invalid" in info
    assert info["winner"].tolist() == [0] * 5 and info["reward_p1"].tolist() == [0.0] * 5
    assert info["action_mask_bool"].shape == (5, 36)
    assert torch.equal(info["action_mask_bool"], env.action_mask())
    assert env.turn().tolist() == [0, 0, 0, 1, 1]


def test_pack_unpack_helpers_match_step(cuda):
    import torch
    import qtttgym_b200 as Q
    n = 10_000
    gen, a, b = Q.BatchedEnv(n, seed=4), Q.BatchedEnv(n, seed=4), Q.BatchedEnv(n, seed=4)
    res = torch.empty(n, dtype=torch.int16).pin_memory()
    for ply in range(9):
        info = gen.step_random(record=True)[4]
        act, coin = info["action"].clone(), info["coin"].clone()
        _, r, t, _, i2 = a.step(act, coin)
        b.step_host_packed(Q.pack_actions(act, coin).cpu().pin_memory(), res)
        torch.cuda.synchronize()
        r2, t2, m2, s2 = Q.unpack_result(res)
        assert torch.equal(r.cpu().view(torch.int32), r2.view(torch.int32)) and torch.equal(t.cpu(), t2)
        assert torch.equal(i2["action_mask"].cpu(), m2) and torch.equal(i2["status"].cpu(), s2)
        assert torch.equal(a.state, b.state)


def test_episode_cuda_graph(cuda):
    """A whole episode (reset + 9 steps) captured in one CUDA graph replays to the same states."""
    import torch
    import qtttgym_b200 as Q
    n = 4096
    gen = Q.BatchedEnv(n, seed=21)
    acts, coins = [], []
    for _ in range(9):
        info = gen.step_random(record=True)[4]
        acts.append(info["action"].clone()); coins.append(info["coin"].clone())
    acts, coins = torch.stack(acts), torch.stack(coins)
    env = Q.BatchedEnv(n, seed=21)
    graph = env.capture_episode(acts, coins)
    for _ in range(3):
        env.state.fill_(-1)                  # garbage: the graph's reset must overwrite it
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(env.state, gen.state)
        assert torch.equal(env.done, gen.done) and torch.equal(env.mask, gen.mask)


def test_golden_features(cuda):
    S.check_golden_features(cuda)


def test_golden_getmask(cuda):
    S.check_golden_getmask(cuda)


@pytest.mark.parametrize("n", [1, 33, 5000, 70_001])
def test_fused_step_features(cuda, n):
    """qttt_step_features: ragged sizes (partial warps, many blocks)"""
    S.check_step_features(cuda, n)


@pytest.mark.parametrize("n", [1, 255, 3000, 70_001])
def test_fused_step_obs(cuda, n):
    """qttt_step_obs: step + the env.py observation from one launch, ragged sizes, all modes"""
    S.check_step_obs(cuda, n)


def test_full_obs_mode_and_vector_env_use_the_fused_step(cuda):
    """BatchedEnv(obs_mode='full').step and VectorEnv.step return the observation of the new state
    (qttt_step_obs) -- the same tensors a separate observation() call decodes."""
    import torch
    import qtttgym_b200 as Q
    env = Q.BatchedEnv(1000, seed=2, obs_mode="full")
    env.reset()
    for _ in range(7):
        masks = env.action_mask().float()
        masks[masks.sum(1) == 0, 0] = 1.0
        act = torch.multinomial(masks, 1).squeeze(1).to(torch.uint8)
        obs, _, _, _, _ = env.step(act)
        want = env.observation()
        for k in want:
            assert torch.equal(obs[k], want[k]), k
    # fresh=True: reset + first step + observation in one launch == reset_step, then observation()
    a, b = Q.BatchedEnv(1000, seed=5), Q.BatchedEnv(1000, seed=5)
    for e in (a, b):
        for _ in range(5):
            e.step_random()
    act = torch.randint(0, 36, (1000,), dtype=torch.uint8, device="cuda")
    obs, r1, d1, _, i1 = a.step_obs(act, fresh=True)
    _, r2, d2, _, i2 = b.reset_step(act)
    assert torch.equal(a.state, b.state) and torch.equal(r1, r2) and torch.equal(d1, d2)
    assert torch.equal(i1["action_mask"], i2["action_mask"]) and a.epoch == b.epoch
    want = b.observation()
    for k in want:
        assert torch.equal(obs[k], want[k]), k
    venv = Q.VectorEnv(777, seed=3)
    venv.reset()
    for _ in range(15):
        obs, _, _, _, _ = venv.step(venv.sample_actions())
        want = venv.env.observation()
        for k in want:
            assert torch.equal(obs[k], want[k]), k


def test_mcts_many_roots_per_block_equals_one_root_per_block(cuda):
    """With more roots than one-warp blocks keep resident, qttt_mcts_run packs several roots into a
    block: the same trees, visit for visit, as the same roots searched 1000 at a time (each shard
    keyed by its root_base, one root per block)."""
    import torch
    import qtttgym_b200 as Q
    n = 4000
    env = Q.BatchedEnv(n, seed=11)
    for _ in range(4):
        env.step_random()
    roots = env.state.clone()
    big = Q.BatchedMCTS(rollouts=60, num_simulations=10, seed=5).reset(roots)
    big.contemplate()
    n_big, q_big, tot_big = big.root_stats()[:3]
    pick_big = big.choose()
    assert int(big.errors().max().item()) == 0
    for lo in range(0, n, 1000):
        part = Q.BatchedMCTS(rollouts=60, num_simulations=10, seed=5, root_base=lo).reset(roots[lo:lo + 1000])
        part.contemplate()
        n_p, q_p, tot_p = part.root_stats()[:3]
        assert torch.equal(n_p, n_big[lo:lo + 1000]) and torch.equal(q_p, q_big[lo:lo + 1000])
        assert torch.equal(tot_p, tot_big[lo:lo + 1000]) and torch.equal(part.choose(), pick_big[lo:lo + 1000])


def test_batches_beyond_2_31_games_are_sliced_correctly(cuda):
    """The kernels index games with 32 bits; the entry points cut larger batches into slices of
    2^31 (2^30 for the boards-only qeval).  2^31 + 70,001 games (34 GB of states) stepped three plies
    through qttt_step_ex, then qttt_qeval_both, qttt_step_packed and qttt_step_obs: the games on both sides of the
    slice boundaries and at the very end equal the same games processed as a small batch."""
    import torch
    import qtttgym_b200 as Q
    import qtttgym_b200._lib as L
    n = (1 << 31) + 70_001
    free, _ = torch.cuda.mem_get_info()
    if free < 120 * (1 << 30):
        pytest.skip("needs 120 GB of free device memory")
    lib = L.lib()
    dev = torch.device("cuda")
    state = torch.empty((n, 4), dtype=torch.int32, device=dev)
    mask = torch.empty(n, dtype=torch.int64, device=dev)
    done = torch.empty(n, dtype=torch.uint8, device=dev)
    plies = [(torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))
             for _ in range(3)]
    step = 1 << 27                                                   # generated in pieces: small temporaries
    for lo in range(0, n, step):
        idx = torch.arange(lo, min(n, lo + step), device=dev, dtype=torch.int64)
        for ply, (act, coin) in enumerate(plies):
            act[lo:lo + step] = ((idx * 7 + ply * 11) % 36).to(torch.uint8)   # some legal, some not; pairs repeat
            coin[lo:lo + step] = ((idx >> 3) & 1).to(torch.uint8)
        del idx
    stream = torch.cuda.current_stream().cuda_stream
    for ply, (act, coin) in enumerate(plies):
        L.check(lib.qttt_step_ex(state.data_ptr(), act.data_ptr(), 0, coin.data_ptr(), 5, 0, 1,
                                 L.STEP_FRESH if ply == 0 else 0, None, done.data_ptr(), mask.data_ptr(), None,
                                 n, stream))
    torch.cuda.synchronize()
    windows = [(0, 1000), ((1 << 31) - 600, (1 << 31) + 600), (n - 1000, n)]
    for lo, hi in windows:
        m = hi - lo
        small = Q.BatchedEnv(m, seed=5, game_base=lo)
        for ply, (act, coin) in enumerate(plies):
            (small.reset_step if ply == 0 else small.step)(act[lo:hi].clone(), coin[lo:hi].clone())
        assert torch.equal(small.state, state[lo:hi]), (lo, hi)
        assert torch.equal(small.mask, mask[lo:hi]) and torch.equal(small.done.to(torch.uint8), done[lo:hi]), (lo, hi)
    # every game of the big batch made progress or was refused: no slice was skipped
    for lo in range(0, n, step):
        nm = (state[lo:lo + step, 0] >> 27) & 15
        assert int(nm.min()) >= 1 and int(nm.max()) <= 3, lo
    del mask, done, nm
    # the boards-only qeval kernel (slices of 2^30) on the same states
    act3 = plies[0][0]
    b0 = torch.empty(n, dtype=torch.int64, device=dev)
    b1 = torch.empty(n, dtype=torch.int64, device=dev)
    closes = torch.empty(n, dtype=torch.uint8, device=dev)
    L.check(lib.qttt_qeval_both(state.data_ptr(), act3.data_ptr(), None, None, b0.data_ptr(), b1.data_ptr(),
                                None, None, closes.data_ptr(), None, n, stream))
    torch.cuda.synchronize()
    for lo, hi in windows + [((1 << 30) - 300, (1 << 30) + 300)]:
        want = Q.qeval_both(state[lo:hi].clone(), act3[lo:hi].clone(), want_states=False, want_probs=False)
        assert torch.equal(want["board0"].view(torch.int64), b0[lo:hi]), (lo, hi)
        assert torch.equal(want["board1"].view(torch.int64), b1[lo:hi]) and torch.equal(want["closes"], closes[lo:hi])
    del b0, b1, closes
    # the packed step (1 byte in, one word out) over the slice boundary
    ac = (plies[1][0] % 36) | (plies[1][1] << 7)
    res = torch.empty(n, dtype=torch.int16, device=dev)
    before = {w: state[w[0]:w[1]].clone() for w in windows}
    L.check(lib.qttt_step_packed(state.data_ptr(), ac.data_ptr(), res.data_ptr(), n, stream))
    torch.cuda.synchronize()
    for (lo, hi), st in before.items():
        r2 = torch.empty(hi - lo, dtype=torch.int16, device=dev)
        L.check(lib.qttt_step_packed(st.data_ptr(), ac[lo:hi].clone().data_ptr(), r2.data_ptr(), hi - lo, stream))
        torch.cuda.synchronize()
        assert torch.equal(st, state[lo:hi]) and torch.equal(r2, res[lo:hi]), (lo, hi)
    del res, ac
    # the fused step + observation launch over the slice boundary
    act, coin = plies[2]
    before = {w: state[w[0]:w[1]].clone() for w in windows}
    cl = torch.empty((n, 9), dtype=torch.int8, device=dev)
    q1 = torch.empty((n, 5, 2), dtype=torch.int8, device=dev)
    q2 = torch.empty((n, 4, 2), dtype=torch.int8, device=dev)
    turn = torch.empty(n, dtype=torch.uint8, device=dev)
    L.check(lib.qttt_step_obs(state.data_ptr(), act.data_ptr(), 0, coin.data_ptr(), 5, 0, 1, 0, None, None, None,
                              None, cl.data_ptr(), q1.data_ptr(), q2.data_ptr(), turn.data_ptr(), n, stream))
    torch.cuda.synchronize()
    for (lo, hi), st in before.items():
        small = Q.BatchedEnv(hi - lo, seed=5, game_base=lo)
        small.state.copy_(st)
        obs = small.step_obs(act[lo:hi].clone(), coin[lo:hi].clone())[0]
        assert torch.equal(small.state, state[lo:hi]), (lo, hi)
        assert torch.equal(obs["classical"], cl[lo:hi]) and torch.equal(obs["q_states_p1"], q1[lo:hi]), (lo, hi)
        assert torch.equal(obs["q_states_p2"], q2[lo:hi]) and torch.equal(obs["turn"], turn[lo:hi]), (lo, hi)
    del state, plies, before, cl, q1, q2, turn
    torch.cuda.empty_cache()


def test_render_states_matches_the_recorded_reference_display(cuda):
    """render_states / BatchedEnv.render on packed states == the displayBoard text recorded from
    the live reference (golden feature records: mid-game, collapsed and autofilled positions)."""
    import numpy as np
    import qtttgym_b200 as Q
    from helpers import load_golden
    recs = load_golden("features_v1.json.gz")
    n = len(recs)
    classical = np.array([r["board"] for r in recs], np.int8)
    moves = np.full((n, 9, 2), -1, np.int8)
    nm = np.array([len(r["moves"]) for r in recs], np.uint8)
    for g, r in enumerate(recs):
        for a, b, idx in r["moves"]:
            moves[g, idx] = (a, b)
    state = Q.pack_states(classical, moves, nm)
    texts = Q.render_states(state)
    assert len(texts) == n
    for g, r in enumerate(recs):
        # displayBoard prints the string followed by print()'s own newline
        assert texts[g] + "\n" == r["display"], g
    env = Q.BatchedEnv(n, seed=0)
    env.state.copy_(state)
    assert env.render(n - 1) + "\n" == recs[-1]["display"]
    assert Q.render_states(state, [3, 1]) == [texts[3], texts[1]]


def test_features_ragged_and_large(cuda):
    """to_vector on ragged sizes vs the host formula applied to observe() output."""
    import torch
    import qtttgym_b200 as Q
    for n in (1, 33, 1000, 70_001):
        env = Q.BatchedEnv(n, seed=n)
        for _ in range(6):
            env.step_random()
        f = Q.to_vector(env.state).cpu().numpy()
        o = {k: v.cpu().numpy() for k, v in env.observation(extras=True).items()}
        want = np.zeros((n, 18, 10), np.float32)
        idx = np.arange(n)
        for sq in range(9):
            col = np.where(o["classical"][:, sq] < 0, 9, o["classical"][:, sq])
            want[idx, sq, col] = 1.0
        live = np.zeros((n, 9), bool)
        for t in range(9):
            has = o["n_moves"] > t
            a, b = o["moves"][:, t, 0], o["moves"][:, t, 1]
            want[idx[has], 9 + a[has], t] = np.float32(1 / 3)
            want[idx[has], 9 + b[has], t] = np.float32(1 / 3)
            uncollapsed = has & (o["classical"][idx, np.maximum(a, 0)] < 0)
            live[idx[uncollapsed], a[uncollapsed]] = True
            live[idx[uncollapsed], b[uncollapsed]] = True
        want[:, 9:, 9] = ~live
        assert np.array_equal(f, want), n


def test_arena_random_vs_random_matches_reference_population(cuda):
    """strat_eval tally convention: random vs random reproduces the reference's win/draw rates."""
    import qtttgym_b200 as Q
    ref = load_golden("population_v1.json")
    n = 200_000
    _, winner = Q.play_games(Q.RandomStrategy(1), Q.RandomStrategy(2), n, seed=3)
    w = np.bincount(winner.cpu().numpy(), minlength=3)
    m = ref["games"]
    for got, key in ((w[1], "x"), (w[2], "o"), (w[0], "draw")):
        p, q = got / n, ref[key] / m
        assert abs(p - q) < 5 * (q * (1 - q) * (1 / n + 1 / m)) ** 0.5, (key, p, q)


def test_arena_rollout_player_beats_random(cuda):
    import qtttgym_b200 as Q
    res = Q.eval_strats(Q.RolloutStrategy(n_rollouts=24, seed=5), Q.RandomStrategy(7), num_games=4000, seed=9)
    assert res["games"] == 4000 and res["strat1_wins"] + res["strat2_wins"] + res["draws"] == 4000
    assert res["strat1_wins"] > 0.8 * 4000, res


def test_config5_full_size_sweep_properties(cuda):
    """config 5 at BASELINE.json's full size (1.25e8 games, ~1.04e9 env-steps): size-independent
    properties -- tallies partition the games, the length histogram accounts for every game and
    every step, an 8-way shard split adds up bit-exactly (the multi-GPU invariant), and the
    rates match the reference's population."""
    import qtttgym_b200 as Q
    g, seed = 125_000_000, 20261018
    full = Q.selfplay_sweep(0, g, seed).cpu().numpy()
    assert full[0] + full[1] + full[2] == g == full[5]
    assert full[6:].sum() == g and (full[6:] * np.arange(10)).sum() == full[3]
    assert full[6:11].sum() == 0                      # no game ends before its 5th ply
    assert 1.0e9 < full[3] < 1.1e9
    parts = sum(Q.selfplay_sweep(*Q.shard_range(g, r, 8), seed).cpu().numpy() for r in range(8))
    assert parts.tolist() == full.tolist()
    ref = load_golden("population_v1.json")
    for k, key in enumerate(("x", "o", "draw")):
        q = ref[key] / ref["games"]
        assert abs(full[k] / g - q) < 5 * (q * (1 - q) / ref["games"]) ** 0.5


def test_golden_mcts_search(cuda):
    S.check_golden_mcts_search(cuda)


def test_mcts_batch_vs_oracle(cuda):
    S.check_mcts_batch_vs_oracle(cuda, n_roots=64, rollouts=150, sims=10)
    S.check_mcts_batch_vs_oracle(cuda, n_roots=8, rollouts=40, sims=300, seed=7, root_base=0)


def test_mcts_pool_exhaustion_is_flagged(cuda):
    import torch
    import qtttgym_b200 as Q
    env = Q.BatchedEnv(4)
    m = Q.BatchedMCTS(rollouts=10, num_simulations=4, seed=1)
    m.reset(env.state, total_rollouts=10)
    m.contemplate(400)                       # far beyond the pool
    assert bool((m.errors() & 1).all()) and bool((m.node_counts() <= m.capacity).all())


def test_mcts_player_beats_random(cuda):
    """strat_eval.py:34-95 with the searched player: MCTS (150 rollouts x 8 playouts per move,
    re-rooted with sync() after every ply) against the random policy, both colours."""
    import qtttgym_b200 as Q
    res = Q.eval_strats(Q.MCTSStrategy(rollouts=150, num_simulations=8, seed=5), Q.RandomStrategy(3),
                        num_games=1024, seed=11)
    assert res["strat1_wins"] > 0.72 * res["games"], res     # random-vs-random: X wins 58 %, O 29 %


def test_rollout_frequencies_vs_reference_simulate(cuda):
    S.check_rollout_frequencies_vs_reference_simulate(cuda, n_rollouts=65536)


def test_checkpoint_resume(cuda, tmp_path):
    """A batch saved mid-game and restored into a fresh env continues bit-identically
    (Philox coins are keyed on (seed, game, ply), so no RNG state needs saving)."""
    import torch
    import qtttgym_b200 as Q
    n = 5000
    a = Q.BatchedEnv(n, seed=77, game_base=1000)
    for _ in range(4):
        a.step_random()
    torch.save(a.state_dict(), tmp_path / "ckpt.pt")
    b = Q.BatchedEnv(n, seed=0)
    b.load_state_dict(torch.load(tmp_path / "ckpt.pt"))
    assert torch.equal(a.done, b.done) and torch.equal(a.mask, b.mask)
    assert torch.equal(a.reward.view(torch.int32), b.reward.view(torch.int32))
    for _ in range(5):
        a.step_random()
        b.step_random()
        assert torch.equal(a.state, b.state)


def test_batched_env_api_surface(cuda):
    """reset / step argument forms, obs modes, turn(), action_mask(), invalid()."""
    import torch
    import qtttgym_b200 as Q
    env = Q.BatchedEnv(6, obs_mode="full", seed=1)
    obs, info = env.reset(seed=99, options={"ignored": True})            # Q4: accepted and ignored
    assert set(obs) == {"classical", "q_states_p1", "q_states_p2", "turn"}
    assert bool((obs["classical"] == -1).all()) and int(info["action_mask"][0]) == (1 << 36) - 1
    # python lists of pairs (any order), an illegal one, and an off-board one
    obs, r, term, trunc, info = env.step([[0, 1], [1, 0], [3, 3], [8, 2], [9, 1], [-1, 4]])
    assert env.invalid().tolist() == [False, False, True, False, True, True]
    assert env.turn().tolist() == [1, 1, 0, 1, 0, 0] and obs["turn"].tolist() == [1, 1, 0, 1, 0, 0]
    assert obs["q_states_p1"][0, 0].tolist() == [0, 1] and obs["q_states_p1"][1, 0].tolist() == [0, 1]
    assert obs["q_states_p1"][3, 0].tolist() == [2, 8]
    assert not bool(term.any()) and not bool(trunc.any())
    assert r.view(torch.int32).tolist() == [-2147483648] * 6               # -0.0 everywhere (Q1)
    assert env.action_mask().shape == (6, 36) and bool(env.action_mask().all())
    # index form with wrong batch size / wrong device is refused
    with pytest.raises(ValueError):
        env.step(torch.zeros(5, dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        env.step(torch.zeros(6, dtype=torch.uint8))
    packed = Q.BatchedEnv(6).reset()[0]
    assert set(packed) == {"packed"} and packed["packed"].shape == (6, 4)


def test_reset_step_equals_reset_then_step(cuda):
    import torch
    import qtttgym_b200 as Q
    n = 30_011
    a, b = Q.BatchedEnv(n, seed=2), Q.BatchedEnv(n, seed=2)
    for env in (a, b):                       # leave garbage mid-game states behind
        for _ in range(5):
            env.step_random()
    act = torch.randint(0, 40, (n,), device="cuda").to(torch.uint8)      # some illegal indices
    coin = torch.randint(0, 2, (n,), device="cuda").to(torch.uint8)
    a.reset()
    ra = a.step(act, coin)
    rb = b.reset_step(act, coin)
    assert torch.equal(a.state, b.state)
    assert torch.equal(ra[1].view(torch.int32), rb[1].view(torch.int32)) and torch.equal(ra[2], rb[2])
    assert torch.equal(ra[4]["action_mask"], rb[4]["action_mask"]) and torch.equal(ra[4]["status"], rb[4]["status"])


def test_integration_md_stub_runs_verbatim(cuda):
    """The ctypes stub printed in INTEGRATION.md section 1 is executable as written."""
    import re
    import torch
    import qtttgym_b200._lib as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"## 1\..*?```python\n(.*?)```", text, re.S).group(1)
    block = block.replace('C.CDLL("libqttt_b200.so")', f'C.CDLL("{L.LIB}")')
    ns = {}
    exec(block, ns)
    env = ns["VecEnv"](8, seed=3)
    env.reset()
    pairs = torch.tensor([[0, 1], [1, 0], [3, 3], [2, 8], [9, 1], [4, 5], [7, 6], [0, 8]], dtype=torch.int8, device="cuda")
    reward, done, mask, status = env.step(pairs)
    torch.cuda.synchronize()
    assert status.tolist() == [0, 0, 1, 0, 1, 0, 0, 0]
    assert reward.view(torch.int32).tolist() == [-2147483648] * 8 and not bool(done.any())
    assert int(mask[0]) == (1 << 36) - 1
