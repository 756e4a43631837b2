"""Two ways to run the packed-state transition, behind one numpy-facing interface:

* ``CudaBackend``  -- the product: qtttgym_b200 (ctypes -> libqttt_b200.so -> CUDA kernels).
* ``EmuBackend``   -- TEST ONLY: tests/hostemu/hostemu.cpp compiles the very same per-game
  functions (qttt_core.cuh) for the CPU so the logic can be checked without a GPU.

Both expose what oracle.c_oracle.Games exposes, so parity_suite.py can diff any of them
against the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------ host emulation
class EmuBackend:
    name = "hostemu"
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            src = os.path.join(HERE, "hostemu", "hostemu.cpp")
            core = os.path.join(ROOT, "qtttgym_b200", "csrc", "qttt_core.cuh")
            out_dir = os.path.join(HERE, "hostemu", "_build")
            os.makedirs(out_dir, exist_ok=True)
            out = os.path.join(out_dir, "libqttt_hostemu.so")
            mcts = os.path.join(ROOT, "qtttgym_b200", "csrc", "qttt_mcts.cuh")
            if (not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(core),
                                                                        os.path.getmtime(mcts))):
                subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                                "-o", out, src], check=True)
            cls._lib = C.CDLL(out)
        return cls._lib

    def games(self, n):
        return EmuGames(n)

    def sweep(self, lo, hi, seed):
        stats = np.zeros(16, np.int64)
        self.lib().emu_sweep(C.c_int64(lo), C.c_int64(hi), C.c_uint64(seed), _p(stats))
        return stats

    def mcts(self, states, num_simulations, seed, root_base, budget):
        return EmuMCTS(self.lib(), states, num_simulations, seed, root_base, budget)


class EmuMCTS:
    def __init__(self, lib, states, num_simulations, seed, root_base, budget):
        self.lib, self.sims, self.seed, self.base = lib, num_simulations, seed, root_base
        self.states = np.ascontiguousarray(states, np.uint32).reshape(-1, 4)
        self.n = self.states.shape[0]
        self.cap = 1 + 2 * budget + 18
        self.pool = np.zeros(self.n * self.cap * lib.emu_mcts_node_bytes(), np.uint8)
        self.meta = np.zeros((self.n, 8), np.int32)
        lib.emu_mcts_init(_p(self.pool), C.c_int64(self.cap), _p(self.meta), _p(self.states), C.c_int64(self.n))

    def contemplate(self, n_rollouts):
        self.lib.emu_mcts_run(_p(self.pool), C.c_int64(self.cap), _p(self.meta), C.c_int32(n_rollouts),
                              C.c_int32(self.sims), C.c_double(1.0), C.c_uint64(self.seed),
                              C.c_uint64(self.base), C.c_int64(self.n))

    def stats(self):
        n = np.zeros((self.n, 36), np.int32); q = np.zeros((self.n, 36), np.float64)
        ntot = np.zeros(self.n, np.int32); ch = np.zeros(self.n, np.uint8)
        self.lib.emu_mcts_stats(_p(self.pool), C.c_int64(self.cap), _p(self.meta), _p(n), _p(q), _p(ntot),
                                _p(ch), C.c_int64(self.n))
        return n, q, ntot, ch

    def sync(self, actions, states):
        ac = np.ascontiguousarray(actions, np.uint8)
        st = np.ascontiguousarray(states, np.uint32).reshape(-1, 4)
        self.lib.emu_mcts_sync(_p(self.pool), C.c_int64(self.cap), _p(self.meta), _p(ac), _p(st), C.c_int64(self.n))

    def errors(self):
        return self.meta[:, 3].copy()

    def taken(self):
        return self.meta[:, 1].copy()

    def live(self):
        return (self.meta[:, 1] - self.meta[:, 5]).copy()


class EmuGames:
    def __init__(self, n):
        self.n = int(n)
        self.state = np.zeros((self.n, 4), np.uint32)
        self.lib = EmuBackend.lib()
        self.lib.emu_reset(_p(self.state), None, C.c_int64(self.n))

    def _outs(self):
        n = self.n
        return dict(reward=np.empty(n, np.float32), done=np.empty(n, np.uint8),
                    mask=np.empty(n, np.uint64), status=np.empty(n, np.uint8))

    def step(self, pairs, coins=None):
        ap = np.ascontiguousarray(pairs, dtype=np.int8).reshape(self.n, 2)
        co = None if coins is None else np.ascontiguousarray(coins, dtype=np.uint8)
        o = self._outs()
        self.lib.emu_step(_p(self.state), _p(ap), 1, _p(co), C.c_uint64(0), C.c_uint64(0),
                          _p(o["reward"]), _p(o["done"]), _p(o["mask"]), _p(o["status"]),
                          C.c_int64(self.n))
        return o

    def step_index(self, actions, coins=None, seed=0, game_base=0):
        ac = np.ascontiguousarray(actions, dtype=np.uint8)
        co = None if coins is None else np.ascontiguousarray(coins, dtype=np.uint8)
        o = self._outs()
        self.lib.emu_step(_p(self.state), _p(ac), 0, _p(co), C.c_uint64(seed), C.c_uint64(game_base),
                          _p(o["reward"]), _p(o["done"]), _p(o["mask"]), _p(o["status"]),
                          C.c_int64(self.n))
        return o

    def step_ex(self, actions, coins=None, seed=0, game_base=0, epoch=0, flags=0):
        ac = np.ascontiguousarray(actions, dtype=np.uint8)
        co = None if coins is None else np.ascontiguousarray(coins, dtype=np.uint8)
        o = self._outs()
        rc = self.lib.emu_step_ex(_p(self.state), _p(ac), 0, _p(co), C.c_uint64(seed), C.c_uint64(game_base),
                                  C.c_uint64(epoch), C.c_uint32(flags), _p(o["reward"]), _p(o["done"]),
                                  _p(o["mask"]), _p(o["status"]), C.c_int64(self.n))
        assert rc == 0
        return o

    def step_random_ex(self, seed, game_base=0, epoch=0, flags=0):
        o = self._outs()
        o["action"] = np.empty(self.n, np.uint8)
        o["coin"] = np.empty(self.n, np.uint8)
        rc = self.lib.emu_step_random_ex(_p(self.state), C.c_uint64(seed), C.c_uint64(game_base),
                                         C.c_uint64(epoch), C.c_uint32(flags), _p(o["action"]), _p(o["coin"]),
                                         _p(o["reward"]), _p(o["done"]), _p(o["mask"]), _p(o["status"]),
                                         C.c_int64(self.n))
        assert rc == 0
        return o

    def step_packed(self, action_coin):
        ac = np.ascontiguousarray(action_coin, dtype=np.uint8)
        res = np.empty(self.n, np.uint16)
        self.lib.emu_step_packed(_p(self.state), _p(ac), _p(res), C.c_int64(self.n))
        return res

    def step_random(self, seed, game_base=0):
        o = self._outs()
        o["action"] = np.empty(self.n, np.uint8)
        o["coin"] = np.empty(self.n, np.uint8)
        self.lib.emu_step_random(_p(self.state), C.c_uint64(seed), C.c_uint64(game_base),
                                 _p(o["action"]), _p(o["coin"]), _p(o["reward"]), _p(o["done"]),
                                 _p(o["mask"]), _p(o["status"]), C.c_int64(self.n))
        return o

    def observe(self):
        n = self.n
        o = dict(classical=np.empty((n, 9), np.int8), moves=np.empty((n, 9, 2), np.int8),
                 n_moves=np.empty(n, np.uint8), q_p1=np.empty((n, 5, 2), np.int8),
                 q_p2=np.empty((n, 4, 2), np.int8), turn=np.empty(n, np.uint8),
                 rounds=np.empty((n, 2), np.int8), reward_p1=np.empty(n, np.float32),
                 winner=np.empty(n, np.uint8), mask_bool=np.empty((n, 36), np.uint8))
        self.lib.emu_observe(_p(self.state), _p(o["classical"]), _p(o["moves"]), _p(o["n_moves"]),
                             _p(o["q_p1"]), _p(o["q_p2"]), _p(o["turn"]), _p(o["rounds"]),
                             _p(o["reward_p1"]), _p(o["winner"]), _p(o["mask_bool"]), C.c_int64(n))
        return o

    def features(self):
        out = np.empty((self.n, 18, 10), np.float32)
        self.lib.emu_features(_p(self.state), _p(out), C.c_int64(self.n))
        return out

    def get_mask(self):
        out = np.empty((self.n, 36), np.uint8)
        self.lib.emu_get_mask(_p(self.state), _p(out), C.c_int64(self.n))
        return out.astype(bool)

    def step_obs(self, actions, coins=None, epoch=0, flags=0, pairs=False):
        """host emulation: the step, then the plain observation of the new state"""
        if pairs:
            assert epoch == 0 and flags == 0
            o = self.step(actions, coins)
        else:
            o = self.step_ex(actions, coins, epoch=epoch, flags=flags)
        obs = self.observe()
        o["obs"] = {k: obs[k] for k in ("classical", "q_p1", "q_p2", "turn")}
        return o

    def step_features(self, actions, coins=None, epoch=0, flags=0):
        """host emulation: the step, then the two encoders on the new state"""
        o = self.step_ex(actions, coins, epoch=epoch, flags=flags)
        o["features"], o["illegal_mask"] = self.features(), self.get_mask()
        return o

    def load(self, classical, moves, n_moves):
        cl = np.ascontiguousarray(classical, np.int8)
        mv = np.ascontiguousarray(moves, np.int8)
        nm = np.ascontiguousarray(n_moves, np.uint8)
        self.lib.emu_pack(_p(self.state), _p(cl), _p(mv), _p(nm), C.c_int64(self.n))
        return self

    def qeval_both(self, actions, squares=True, boards_only=False):
        """squares=False takes the one-sweep path (no per-move squares), like the library."""
        squares = squares and not boards_only
        n = self.n
        ac = np.ascontiguousarray(actions, np.uint8)
        o = dict(next0=np.empty((n, 4), np.uint32), next1=np.empty((n, 4), np.uint32),
                 board0=np.empty(n, np.uint64), board1=np.empty(n, np.uint64),
                 closes=np.empty(n, np.uint8), result_prob=np.empty((n, 3), np.float32))
        if boards_only:
            o = {k: o[k] for k in ("board0", "board1", "closes")}
        if squares:
            o.update(sq0=np.empty((n, 9), np.int8), sq1=np.empty((n, 9), np.int8))
        self.lib.emu_qeval_both(_p(self.state), _p(ac), _p(o.get("next0")), _p(o.get("next1")),
                                _p(o["board0"]), _p(o["board1"]), _p(o.get("sq0")), _p(o.get("sq1")),
                                _p(o["closes"]), _p(o.get("result_prob")), C.c_int64(n))
        return o

    def rollout(self, n_rollouts, seed):
        tallies = np.empty((self.n, 3), np.int32)
        value = np.empty(self.n, np.float32)
        steps = np.zeros(1, np.int64)
        self.lib.emu_rollout(_p(self.state), C.c_int64(self.n), C.c_int32(n_rollouts),
                             C.c_uint64(seed), _p(tallies), _p(value), _p(steps))
        return tallies, value, int(steps[0])

    def with_state(self, state):
        g = EmuGames.__new__(EmuGames)
        g.lib = self.lib
        g.state = np.ascontiguousarray(state, np.uint32).reshape(-1, 4).copy()
        g.n = g.state.shape[0]
        return g


# ------------------------------------------------------------------------------ CUDA (product)
class CudaBackend:
    name = "cuda"

    def games(self, n):
        return CudaGames(n)

    def sweep(self, lo, hi, seed):
        import qtttgym_b200 as Q
        return Q.selfplay_sweep(lo, hi, seed).cpu().numpy()

    def mcts(self, states, num_simulations, seed, root_base, budget):
        return CudaMCTS(states, num_simulations, seed, root_base, budget)


class CudaMCTS:
    def __init__(self, states, num_simulations, seed, root_base, budget):
        import torch
        import qtttgym_b200 as Q
        self.torch = torch
        st = torch.from_numpy(np.ascontiguousarray(states).view(np.int32).reshape(-1, 4)).cuda()
        self.m = Q.BatchedMCTS(rollouts=budget, num_simulations=num_simulations, seed=seed, root_base=root_base)
        self.m.reset(st, total_rollouts=budget)

    def contemplate(self, n_rollouts):
        self.m.contemplate(n_rollouts)

    def stats(self):
        n, q, ntot, ch = self.m.root_stats()
        return n.cpu().numpy(), q.cpu().numpy(), ntot.cpu().numpy(), ch.cpu().numpy()

    def sync(self, actions, states):
        t = self.torch
        self.m.sync(t.from_numpy(np.ascontiguousarray(actions, np.uint8)).cuda(),
                    t.from_numpy(np.ascontiguousarray(states).view(np.int32).reshape(-1, 4)).cuda())

    def errors(self):
        return self.m.errors().cpu().numpy()

    def taken(self):
        return self.m.node_counts().cpu().numpy()

    def live(self):
        return self.m.live_counts().cpu().numpy()


class CudaGames:
    def __init__(self, n, seed=0, game_base=0):
        import torch
        import qtttgym_b200 as Q
        self.torch, self.Q = torch, Q
        self.n = int(n)
        self.env = Q.BatchedEnv(self.n, seed=seed, game_base=game_base)

    def _outs(self, res):
        _, reward, term, _, info = res
        t = self.torch
        t.cuda.synchronize()
        return dict(reward=reward.cpu().numpy().copy(), done=term.cpu().numpy().astype(np.uint8),
                    mask=info["action_mask"].cpu().numpy().astype(np.uint64),
                    status=info["status"].cpu().numpy().copy())

    def step(self, pairs, coins=None):
        t = self.torch
        ap = t.from_numpy(np.ascontiguousarray(pairs, dtype=np.int8).reshape(self.n, 2)).cuda()
        co = None if coins is None else t.from_numpy(np.ascontiguousarray(coins, np.uint8)).cuda()
        return self._outs(self.env.step(ap, co))

    def step_index(self, actions, coins=None, seed=0, game_base=0):
        t = self.torch
        self.env.seed, self.env.game_base = seed, game_base
        ac = t.from_numpy(np.ascontiguousarray(actions, np.uint8)).cuda()
        co = None if coins is None else t.from_numpy(np.ascontiguousarray(coins, np.uint8)).cuda()
        return self._outs(self.env.step(ac, co))

    _AUTORESET = {0: False, 2: True, 4: "next"}

    def step_ex(self, actions, coins=None, seed=0, game_base=0, epoch=0, flags=0):
        """through BatchedEnv.step / reset_step (which own the epoch counter: it is preset here
        so that the call lands on the requested epoch)."""
        t = self.torch
        env = self.env
        env.seed, env.game_base = seed, game_base
        ac = t.from_numpy(np.ascontiguousarray(actions, np.uint8)).cuda()
        co = None if coins is None else t.from_numpy(np.ascontiguousarray(coins, np.uint8)).cuda()
        if flags == 1:
            env.epoch = epoch - 1
            return self._outs(env.reset_step(ac, co))
        env.epoch = epoch - (1 if flags else 0)
        return self._outs(env.step(ac, co, autoreset=self._AUTORESET[flags]))

    def step_random_ex(self, seed, game_base=0, epoch=0, flags=0):
        env = self.env
        env.seed, env.game_base = seed, game_base
        env.epoch = epoch - (1 if flags else 0)
        res = env.step_random(record=True, autoreset=self._AUTORESET[flags])
        o = self._outs(res)
        o["action"] = res[4]["action"].cpu().numpy()
        o["coin"] = res[4]["coin"].cpu().numpy()
        return o

    def step_packed(self, action_coin, variant="copy"):
        """variant: "copy" (cudaMemcpyAsync pipeline), "copy_obs" (+ observation), "mapped"
        (the kernel reads/writes the pinned host buffers itself), "mapped_obs", "mapped12" (bit-packed results)."""
        t = self.torch
        ac = t.from_numpy(np.ascontiguousarray(action_coin, np.uint8)).pin_memory()
        if variant in ("mapped12", "copy12"):        # 12-bit results, four envs in three words
            res12 = t.full((3 * ((self.n + 3) // 4),), -1, dtype=t.int16).pin_memory()
            self.env.step_host_packed12(ac, res12, mapped=(variant == "mapped12"), chunks=3, n_streams=2)
            t.cuda.synchronize()
            return self.Q.unpack_result12(res12, self.n).numpy().view(np.uint16).copy()
        res = t.empty(self.n, dtype=t.int16).pin_memory()
        obs = t.empty((self.n, 4), dtype=t.int32).pin_memory() if variant.endswith("obs") else None
        self.env.step_host_packed(ac, res, obs_host=obs, chunks=3, n_streams=2, mapped=variant.startswith("mapped"))
        t.cuda.synchronize()
        if obs is not None:
            self.last_obs = obs.numpy().view(np.uint32).copy()
        return res.numpy().view(np.uint16).copy()

    def step_random(self, seed, game_base=0):
        self.env.seed, self.env.game_base = seed, game_base
        res = self.env.step_random(record=True)
        o = self._outs(res)
        o["action"] = res[4]["action"].cpu().numpy()
        o["coin"] = res[4]["coin"].cpu().numpy()
        return o

    def observe(self):
        o = self.env.observation(extras=True)
        out = {k: v.cpu().numpy() for k, v in o.items()}
        out["q_p1"], out["q_p2"] = out.pop("q_states_p1"), out.pop("q_states_p2")
        out["mask_bool"] = out.pop("action_mask").astype(np.uint8)
        return out

    def features(self):
        return self.Q.to_vector(self.env.state).cpu().numpy()

    def get_mask(self):
        return self.Q.get_mask(self.env.state).cpu().numpy()

    def step_obs(self, actions, coins=None, epoch=0, flags=0, pairs=False):
        """the fused kernel: qttt_step_obs"""
        t = self.torch
        env = self.env
        if pairs:
            ac = t.from_numpy(np.ascontiguousarray(actions, dtype=np.int8).reshape(self.n, 2)).cuda()
        else:
            ac = t.from_numpy(np.ascontiguousarray(actions, np.uint8)).cuda()
        co = None if coins is None else t.from_numpy(np.ascontiguousarray(coins, np.uint8)).cuda()
        env.epoch = epoch - (1 if flags else 0)
        res = env.step_obs(ac, co, autoreset=self._AUTORESET[flags])
        o = self._outs(res)
        obs = {k: v.cpu().numpy() for k, v in res[0].items()}
        o["obs"] = {"classical": obs["classical"], "q_p1": obs["q_states_p1"], "q_p2": obs["q_states_p2"],
                    "turn": obs["turn"]}
        return o

    def step_features(self, actions, coins=None, epoch=0, flags=0):
        """the fused kernel: qttt_step_features"""
        t = self.torch
        env = self.env
        ac = t.from_numpy(np.ascontiguousarray(actions, np.uint8)).cuda()
        co = None if coins is None else t.from_numpy(np.ascontiguousarray(coins, np.uint8)).cuda()
        env.epoch = epoch - (1 if flags else 0)
        res = env.step_features(ac, co, autoreset=self._AUTORESET[flags], want_mask=True)
        o = self._outs(res)
        o["features"] = res[4]["features"].cpu().numpy()
        o["illegal_mask"] = res[4]["illegal_mask"].cpu().numpy()
        return o

    def load(self, classical, moves, n_moves):
        self.env.load_positions(np.asarray(classical, np.int8), np.asarray(moves, np.int8),
                                np.asarray(n_moves, np.uint8))
        return self

    @property
    def state(self):
        return self.env.state.cpu().numpy().view(np.uint32)

    def qeval_both(self, actions, squares=True, boards_only=False):
        t = self.torch
        ac = t.from_numpy(np.ascontiguousarray(actions, np.uint8)).cuda()
        if boards_only:      # the config-3 shape: its own kernel (k_qeval_boards)
            res = self.Q.qeval_both(self.env.state, ac, want_states=False, want_probs=False)
        else:
            res = self.Q.qeval_both(self.env.state, ac, want_squares=squares)
        out = {k: v.cpu().numpy() for k, v in res.items()}
        for k in ("next0", "next1"):
            if k in out:
                out[k] = out[k].view(np.uint32)
        for k in ("board0", "board1"):
            out[k] = out[k].astype(np.uint64)
        return out

    def rollout(self, n_rollouts, seed):
        tallies, value, steps = self.Q.rollout_eval(self.env.state, n_rollouts, seed)
        return tallies.cpu().numpy(), value.cpu().numpy(), int(steps.item())

    def with_state(self, state):
        t = self.torch
        g = CudaGames(np.asarray(state).reshape(-1, 4).shape[0])
        g.env.state.copy_(t.from_numpy(np.ascontiguousarray(state).view(np.int32).reshape(-1, 4)).cuda())
        return g
