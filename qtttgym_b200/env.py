"""Batched, GPU-resident stand-in for ``qtttgym.Env`` (reference: qtttgym/env.py:15-112).

``BatchedEnv`` holds N independent games as packed 16-byte states in HBM and advances all of
them with one kernel launch per ``step``.  The call shapes follow the reference's gym-like
API -- ``reset() -> (obs, info)``, ``step(action) -> (obs, reward, terminated, truncated,
info)`` -- with tensors in place of Python scalars.  ``Env`` is the ``num_envs == 1`` adapter
that reproduces the reference's exact return types so it drops into ``qtttgym.Env`` call
sites.  All arithmetic happens in libqttt_b200.so (CUDA, sm_100a); there is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib
from .actions import PAIRS


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class BatchedEnv:
    """N quantum tic-tac-toe games on one GPU.

    Parameters
    ----------
    num_envs : number of games.
    device   : CUDA device.
    seed     : Philox key for collapse coins (when ``choices`` is not forced) and for
               ``step_random``; draws are keyed ``(seed, game_base + env index, len(moves))``
               so results do not depend on how games are sharded over GPUs.
    game_base: global id of env 0 (rank offset in multi-GPU runs).
    obs_mode : ``"packed"`` (default) -> ``obs = {"packed": int32[N,4]}``, the live state tensor
               (like the reference, whose obs aliases the live board -- quirk Q5);
               ``"full"`` -> snapshot tensors per ``observation()``.
    """

    def __init__(self, num_envs: int, device="cuda", seed: int = 0, game_base: int = 0,
                 obs_mode: str = "packed"):
        if obs_mode not in ("packed", "full"):
            raise ValueError("obs_mode must be 'packed' or 'full'")
        self.lib = _lib.lib()                       # raises if the CUDA library is missing
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("qtttgym_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.game_base = int(game_base)
        self.obs_mode = obs_mode
        n, dev = self.num_envs, self.device
        self.state = torch.zeros((n, 4), dtype=torch.int32, device=dev)
        self.reward = torch.empty(n, dtype=torch.float32, device=dev)
        self.done = torch.empty(n, dtype=torch.bool, device=dev)      # kernel writes 0 / 1 bytes
        self._never = torch.zeros(n, dtype=torch.bool, device=dev)    # ``truncated`` (Q8)
        self.mask = torch.empty(n, dtype=torch.int64, device=dev)     # 36-bit legal mask
        self.status = torch.empty(n, dtype=torch.uint8, device=dev)
        self._host_streams = None            # lazily created by step_host
        self._stream_array = None
        self._d_act = self._d_coin = self._d_res16 = None
        # Episode counter folded into the Philox counter of the collapse coins: bumped by every
        # reset (and by every step of an auto-resetting batch), so that successive episodes in
        # the same env slots do not replay the same coins.  Part of state_dict().
        self.epoch = 0
        self._first_reset = True
        self.reset()

    # ------------------------------------------------------------------ gym-like API
    def reset(self, *, seed=None, options=None):
        """env.py:55-57.  ``seed`` / ``options`` are accepted and ignored, as in the reference (Q4)."""
        if self._first_reset:
            self._first_reset = False         # the constructor's reset: epoch 0 is the first episode
        else:
            self.epoch += 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_reset_all(self.state.data_ptr(), self.mask.data_ptr(),
                                               self.reward.data_ptr(), self.done.data_ptr(),
                                               self.status.data_ptr(), self.num_envs,
                                               _stream_ptr(self.device)))
        return self._obs(), _Info(self, {"action_mask": self.mask})

    @staticmethod
    def _mode_flags(autoreset):
        if not autoreset:
            return 0
        if autoreset is True or autoreset == "apply":
            return _lib.STEP_AUTORESET
        if autoreset == "next":
            return _lib.STEP_AUTORESET_NEXT
        raise ValueError("autoreset must be False, True / 'apply' or 'next'")

    def step(self, actions, choices=None, autoreset=False):
        """env.py:34-53 for every env.

        actions : uint8[N] action indices 0..35 (mcts.py:339-349) **or** int8[N,2] ``(a, b)``
                  pairs as passed to the reference's ``Env.step`` (any order).  Illegal actions
                  are swallowed no-ops exactly as env.py:36-43 (``info["status"] == 1``, ``env.invalid()``).
        choices : uint8[N] forced collapse coins (0 -> the closing move falls into its smaller
                  square, qeval.py:35), consumed only by envs whose move closes a cycle;
                  ``None`` -> Philox coins.
        autoreset : ``False`` -- the reference's behaviour: finished games keep accepting moves
                  (Q3) or idle on illegal ones; ``True`` / ``"apply"`` -- a game that is over on
                  entry restarts from the empty board inside the same launch and the action is
                  applied to the fresh game; ``"next"`` -- it restarts and its action is ignored
                  (the next-step autoreset of vector envs).  ``info["status"] & 4`` marks the envs
                  that were reset.
        Returns ``(obs, reward f32[N], terminated bool[N], truncated bool[N], info)``; reward is
        the reference's -1.0 / -0.0 (quirk Q1).  The returned tensors are views of buffers that
        the next ``step`` overwrites.
        """
        if self.obs_mode == "full":          # the observation (fresh tensors) comes out of the same launch
            return self.step_obs(actions, choices, autoreset)
        act, fmt = self._check_actions(actions)
        coin = self._check_choices(choices)
        flags = self._mode_flags(autoreset)
        if flags:
            self.epoch += 1        # per-env episodes diverge: every step of the batch gets its own epoch
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_step_ex(
                self.state.data_ptr(), act.data_ptr(), fmt, _lib.ptr(coin), self.seed,
                self.game_base, self.epoch, flags, self.reward.data_ptr(), self.done.data_ptr(),
                self.mask.data_ptr(), self.status.data_ptr(), self.num_envs,
                _stream_ptr(self.device)))
        return self._result()

    def step_obs(self, actions, choices=None, autoreset=False, out=None, fresh: bool = False):
        """``step`` fused with the env.py observation of the new states (``qttt_step_obs``): one
        launch steps the games and decodes ``classical`` int8[N,9], ``q_states_p1`` int8[N,5,2],
        ``q_states_p2`` int8[N,4,2], ``turn`` uint8[N] (env.py:68-85) from the states it still holds
        in registers.  ``out``: the observation dict of an earlier call, overwritten in place.
        ``fresh``: ``reset()`` first, inside the same launch (like ``reset_step``).
        Returns what ``step`` returns, with that dict as ``obs``."""
        act, fmt = self._check_actions(actions)
        coin = self._check_choices(choices)
        flags = _lib.STEP_FRESH if fresh else self._mode_flags(autoreset)
        if flags:
            self.epoch += 1
        n, dev = self.num_envs, self.device
        if out is None:
            out = {"classical": torch.empty((n, 9), dtype=torch.int8, device=dev),
                   "q_states_p1": torch.empty((n, 5, 2), dtype=torch.int8, device=dev),
                   "q_states_p2": torch.empty((n, 4, 2), dtype=torch.int8, device=dev),
                   "turn": torch.empty(n, dtype=torch.uint8, device=dev)}
        with torch.cuda.device(dev):
            _lib.check(self.lib.qttt_step_obs(
                self.state.data_ptr(), act.data_ptr(), fmt, _lib.ptr(coin), self.seed,
                self.game_base, self.epoch, flags, self.reward.data_ptr(), self.done.data_ptr(),
                self.mask.data_ptr(), self.status.data_ptr(), out["classical"].data_ptr(),
                out["q_states_p1"].data_ptr(), out["q_states_p2"].data_ptr(), out["turn"].data_ptr(),
                n, _stream_ptr(dev)))
        info = _Info(self, {"action_mask": self.mask, "status": self.status})
        return out, self.reward, self.done, self._never, info

    def _check_choices(self, choices):
        if choices is None:
            return None
        coin = choices if choices.dtype == torch.uint8 else choices.to(torch.uint8)
        coin = coin.contiguous()
        if coin.device != self.device or coin.numel() != self.num_envs:
            raise ValueError("choices must be a uint8[N] tensor on the env's device")
        return coin

    def reset_step(self, actions, choices=None):
        """``reset()`` followed by ``step(actions, choices)`` as ONE kernel launch
        (``qttt_reset_step``): the games restart from the empty board, so the packed state is
        written but never read.  Returns what ``step`` returns."""
        act, fmt = self._check_actions(actions)
        coin = self._check_choices(choices)
        self.epoch += 1                       # a reset: the next episode
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_step_ex(
                self.state.data_ptr(), act.data_ptr(), fmt, _lib.ptr(coin), self.seed,
                self.game_base, self.epoch, _lib.STEP_FRESH, self.reward.data_ptr(),
                self.done.data_ptr(), self.mask.data_ptr(), self.status.data_ptr(), self.num_envs,
                _stream_ptr(self.device)))
        return self._result()

    def step_features(self, actions, choices=None, autoreset=False, want_mask: bool = False, out=None):
        """``step`` fused with the net-input encoding of the new states (``qttt_step_features``):
        ``info["features"]`` float32[N,18,10] is ``GameState.to_vector`` (mcts.py:67-85) of every env
        after the move and, with ``want_mask``, ``info["illegal_mask"]`` bool[N,36] is
        ``nn.Model.get_mask`` (nn.py:44-61) -- written by the launch that steps the games, so the
        policy/value net input costs no second pass over the state array.  ``out``: the ``info`` of
        an earlier call, whose feature / mask tensors are reused."""
        act, fmt = self._check_actions(actions)
        coin = self._check_choices(choices)
        flags = self._mode_flags(autoreset)
        if flags:
            self.epoch += 1
        n, dev = self.num_envs, self.device
        feats = out["features"] if out is not None else torch.empty((n, 18, 10), dtype=torch.float32, device=dev)
        imask = None
        if want_mask:
            imask = out["illegal_mask"] if out is not None and "illegal_mask" in out else \
                torch.empty((n, 36), dtype=torch.bool, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.qttt_step_features(
                self.state.data_ptr(), act.data_ptr(), fmt, _lib.ptr(coin), self.seed, self.game_base,
                self.epoch, flags, self.reward.data_ptr(), self.done.data_ptr(), self.mask.data_ptr(),
                self.status.data_ptr(), feats.data_ptr(), _lib.ptr(imask), n, _stream_ptr(dev)))
        res = self._result()
        res[4]["features"] = feats
        if imask is not None:
            res[4]["illegal_mask"] = imask
        return res

    def step_random(self, record: bool = False, autoreset=False, out=None):
        """One ply of the uniform-random policy of ``MCTS._simulate`` (mcts.py:185-198) for
        every env that is not terminated; terminated envs are left untouched
        (``info["status"] == 2``) unless ``autoreset`` restarts them (see ``step``): then every
        call plays one ply in every env -- continuous random self-play through the step API.
        ``record``: the chosen actions / coins are returned in ``info`` (``out``: a pair of
        uint8[N] tensors to record into)."""
        a_out = c_out = None
        if out is not None:
            a_out, c_out = out
        elif record:
            a_out = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
            c_out = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        flags = self._mode_flags(autoreset)
        if flags:
            self.epoch += 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_step_random_ex(
                self.state.data_ptr(), self.seed, self.game_base, self.epoch, flags,
                _lib.ptr(a_out), _lib.ptr(c_out), self.reward.data_ptr(), self.done.data_ptr(),
                self.mask.data_ptr(), self.status.data_ptr(), self.num_envs,
                _stream_ptr(self.device)))
        res = self._result()
        if a_out is not None:
            res[4]["action"], res[4]["coin"] = a_out, c_out
        return res

    def _host_pipeline(self, n_streams: int):
        """Side streams and device staging buffers of the host-buffer paths; (re)built whenever
        the stream count changes, together with the ctypes array of their handles."""
        import ctypes as C
        if self._host_streams is None or len(self._host_streams) != n_streams:
            n, dev = self.num_envs, self.device
            self._host_streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
            self._stream_array = (C.c_void_p * n_streams)(*[st.cuda_stream for st in self._host_streams])
            if self._d_act is None:
                self._d_act = torch.empty(n, dtype=torch.uint8, device=dev)
                self._d_coin = torch.empty(n, dtype=torch.uint8, device=dev)
                self._d_res16 = torch.empty(n, dtype=torch.int16, device=dev)
        return self._host_streams

    def _slices(self, chunks: int):
        n = self.num_envs
        chunks = max(1, min(chunks, (n + 255) // 256))
        per = -(-n // chunks)
        return -(-per // 256) * 256

    def step_host(self, actions_host, choices_host, reward_host, done_host, mask_host,
                  chunks: int = 8, n_streams: int = 4):
        """``step`` for callers whose buffers live in (pinned) HOST memory -- the end-to-end path.

        actions_host uint8[N] action indices and choices_host uint8[N] coins are copied to the
        device, the step kernel runs, and reward f32[N] / done bool[N] / mask int64[N] are
        copied back into the given host tensors.  The batch is cut into ``chunks`` slices that
        are pipelined over ``n_streams`` side streams so that host->device copies, kernels and
        device->host copies of different slices overlap (PCIe is full duplex).  Returns after
        enqueueing; the current stream waits for all slices, so synchronise it before reading
        the host outputs.
        """
        n, dev = self.num_envs, self.device
        for t, dt in ((actions_host, torch.uint8), (choices_host, torch.uint8),
                      (reward_host, torch.float32), (done_host, torch.bool), (mask_host, torch.int64)):
            if t.dtype != dt or t.numel() != n or t.device.type != "cpu" or not t.is_contiguous():
                raise ValueError("step_host expects contiguous CPU tensors of N elements "
                                 "(uint8 actions, uint8 coins, f32 reward, bool done, int64 mask)")
        streams = self._host_pipeline(n_streams)
        cur = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        per = self._slices(chunks)
        with torch.cuda.device(dev):
            for c in range(-(-n // per)):
                lo, hi = c * per, min(n, (c + 1) * per)
                st = streams[c % n_streams]
                st.wait_event(ready)
                with torch.cuda.stream(st):
                    self._d_act[lo:hi].copy_(actions_host[lo:hi], non_blocking=True)
                    self._d_coin[lo:hi].copy_(choices_host[lo:hi], non_blocking=True)
                    _lib.check(self.lib.qttt_step(
                        self.state.data_ptr() + 16 * lo, self._d_act.data_ptr() + lo, _lib.ACT_INDEX,
                        self._d_coin.data_ptr() + lo, self.seed, self.game_base + lo,
                        self.reward.data_ptr() + 4 * lo, self.done.data_ptr() + lo,
                        self.mask.data_ptr() + 8 * lo, self.status.data_ptr() + lo, hi - lo,
                        st.cuda_stream))
                    reward_host[lo:hi].copy_(self.reward[lo:hi], non_blocking=True)
                    done_host[lo:hi].copy_(self.done[lo:hi], non_blocking=True)
                    mask_host[lo:hi].copy_(self.mask[lo:hi], non_blocking=True)
            for st in streams:
                fin = torch.cuda.Event()
                fin.record(st)
                cur.wait_event(fin)
        return reward_host, done_host, mask_host

    def step_host_packed12(self, action_coin_host, result12_host, mapped: bool = True, chunks: int = 8,
                           n_streams: int = 4):
        """``step_host_packed`` with the result words bit-packed: 12 bits per env, four envs in
        three 16-bit words, so 1.5 instead of 2 bytes per env come back across PCIe -- the direction
        that bounds the host-resident caller.  ``result12_host``: pinned int16[3 * ceil(N / 4)];
        decode with ``unpack_result12(result12_host, N)``.  ``mapped=True``
        (``qttt_step_packed12_mapped``): one launch reads / writes the pinned buffers itself;
        ``mapped=False`` (``qttt_step_packed12_host``): slices pipelined over side streams with
        explicit ``cudaMemcpyAsync`` copies.  Same transition, bit for bit."""
        n, dev = self.num_envs, self.device
        words = 3 * ((n + 3) // 4)
        for t, dt, numel in ((action_coin_host, torch.uint8, n), (result12_host, torch.int16, words)):
            if t.dtype != dt or t.numel() != numel or t.device.type != "cpu" or not t.is_contiguous() or not t.is_pinned():
                raise ValueError("step_host_packed12 expects pinned contiguous CPU uint8[N] / int16[3*ceil(N/4)] tensors")
        if mapped:
            with torch.cuda.device(dev):
                _lib.check(self.lib.qttt_step_packed12_mapped(
                    self.state.data_ptr(), action_coin_host.data_ptr(), result12_host.data_ptr(), n, _stream_ptr(dev)))
            return result12_host
        streams = self._host_pipeline(n_streams)
        cur = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        per = self._slices(chunks)                    # a multiple of 256, hence of 4
        if getattr(self, "_d_res12", None) is None:
            self._d_res12 = torch.empty(words, dtype=torch.int16, device=dev)
        with torch.cuda.device(dev):
            for st in streams:
                st.wait_event(ready)
            _lib.check(self.lib.qttt_step_packed12_host(
                self.state.data_ptr(), action_coin_host.data_ptr(), result12_host.data_ptr(),
                self._d_act.data_ptr(), self._d_res12.data_ptr(), n, per, self._stream_array, n_streams),
                launches=-(-n // per))
            for st in streams:
                fin = torch.cuda.Event()
                fin.record(st)
                cur.wait_event(fin)
        return result12_host

    def step_host_obs12(self, action_coin_host, obs12_host, chunks: int = 8, n_streams: int = 4):
        """``step_host_packed`` returning the OBSERVATION in compact form
        (``qttt_step_packed_host_obs12``): one 12-byte record per env -- the env.py observation
        (classical squares, the uncollapsed moves of both players, hence ``turn``) plus the
        terminated / line / illegal flags -- instead of the 16-byte packed state and a result word.
        ``obs12_host``: pinned int32[N,3]; ``unpack_obs12`` decodes it into what ``step`` and
        ``observation()`` return.  Same transition, bit for bit."""
        n, dev = self.num_envs, self.device
        for t, dt, numel in ((action_coin_host, torch.uint8, n), (obs12_host, torch.int32, 3 * n)):
            if t.dtype != dt or t.numel() != numel or t.device.type != "cpu" or not t.is_contiguous():
                raise ValueError("step_host_obs12 expects contiguous CPU uint8[N] / int32[N,3] tensors")
        streams = self._host_pipeline(n_streams)
        cur = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        per = self._slices(chunks)
        if getattr(self, "_d_obs12", None) is None:
            self._d_obs12 = torch.empty((n, 3), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            for st in streams:
                st.wait_event(ready)
            _lib.check(self.lib.qttt_step_packed_host_obs12(
                self.state.data_ptr(), action_coin_host.data_ptr(), obs12_host.data_ptr(),
                self._d_act.data_ptr(), self._d_obs12.data_ptr(), n, per, self._stream_array, n_streams),
                launches=-(-n // per))
            for st in streams:
                fin = torch.cuda.Event()
                fin.record(st)
                cur.wait_event(fin)
        return obs12_host

    def step_host_packed(self, action_coin_host, result_host, obs_host=None, chunks: int = 8,
                         n_streams: int = 4, mapped: bool = False):
        """``step_host`` with compact I/O: 1 byte in and 2 bytes out per env cross PCIe instead of
        2 + 13.  ``action_coin_host`` uint8[N] from ``pack_actions``; ``result_host`` int16[N],
        decoded with ``unpack_result`` (the 36-bit legal mask is re-expanded on the host from the
        9-bit free-square set).  ``obs_host`` (optional, pinned int32[N,4]) also receives the
        observation -- the packed post-step states, 16 more bytes per env (what the reference's
        ``Env.step`` returns as ``obs``, env.py:46; decode with ``observe_states`` or on the host).
        Same transition, bit for bit.

        ``mapped=False`` (``qttt_step_packed_host[_obs]``): slices pipelined over side streams with
        one ``cudaMemcpyAsync`` per array per slice.  ``mapped=True`` (``qttt_step_packed_mapped``):
        ONE kernel launch on the current stream reads and writes the pinned host buffers itself
        across PCIe (no copy engine, no staging); the buffers must be pinned (``pin_memory()``)."""
        n, dev = self.num_envs, self.device
        checks = [(action_coin_host, torch.uint8, n), (result_host, torch.int16, n)]
        if obs_host is not None:
            checks.append((obs_host, torch.int32, 4 * n))
        for t, dt, numel in checks:
            if t.dtype != dt or t.numel() != numel or t.device.type != "cpu" or not t.is_contiguous():
                raise ValueError("step_host_packed expects contiguous CPU uint8[N] / int16[N] (/ int32[N,4]) tensors")
        if mapped:
            if not (action_coin_host.is_pinned() and result_host.is_pinned()
                    and (obs_host is None or obs_host.is_pinned())):
                raise ValueError("mapped=True needs pinned host tensors (tensor.pin_memory())")
            with torch.cuda.device(dev):
                _lib.check(self.lib.qttt_step_packed_mapped(
                    self.state.data_ptr(), action_coin_host.data_ptr(), result_host.data_ptr(),
                    _lib.ptr(obs_host), n, _stream_ptr(dev)))
            return result_host
        streams = self._host_pipeline(n_streams)
        cur = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        per = self._slices(chunks)
        with torch.cuda.device(dev):
            for st in streams:
                st.wait_event(ready)
            _lib.check(self.lib.qttt_step_packed_host_obs(
                self.state.data_ptr(), action_coin_host.data_ptr(), result_host.data_ptr(),
                _lib.ptr(obs_host), self._d_act.data_ptr(), self._d_res16.data_ptr(), n, per,
                self._stream_array, n_streams), launches=-(-n // per))
            for st in streams:
                fin = torch.cuda.Event()
                fin.record(st)
                cur.wait_event(fin)
        return result_host

    def capture_episode(self, actions, choices):
        """Records ``reset`` + ``len(actions)`` steps (uint8[T,N] action indices and coins, on
        the device) into ONE CUDA graph and returns it; ``graph.replay()`` then plays the whole
        episode with a single launch call -- for small batches, where a step is a few
        microseconds of GPU work and launch latency dominates."""
        if actions.dtype != torch.uint8 or choices.dtype != torch.uint8 or actions.shape != choices.shape \
                or actions.dim() != 2 or actions.shape[1] != self.num_envs or actions.device != self.device:
            raise ValueError("capture_episode expects uint8[T,N] device tensors")
        actions, choices = actions.contiguous(), choices.contiguous()
        self._graph_inputs = (actions, choices)        # keep the captured buffers alive
        self.reset_step(actions[0], choices[0])        # warm-up outside the capture
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.reset_step(actions[0], choices[0])          # reset fused into the first ply
            for t in range(1, actions.shape[0]):
                self.step(actions[t], choices[t])
        return graph

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """Everything needed to resume: the packed states plus the RNG keying.  (The reference
        has no checkpointing of games; here the whole batch is one tensor.)"""
        return {"state": self.state.clone(), "seed": self.seed, "game_base": self.game_base,
                "num_envs": self.num_envs, "epoch": self.epoch}

    def load_state_dict(self, sd):
        if int(sd["num_envs"]) != self.num_envs:
            raise ValueError("checkpoint holds a different number of envs")
        self.state.copy_(sd["state"].to(self.device))
        self.seed, self.game_base = int(sd["seed"]), int(sd["game_base"])
        self.epoch = int(sd.get("epoch", 0))
        # reward / done / mask are functions of the state: recompute them with a no-op step
        noop = torch.full((self.num_envs,), 255, dtype=torch.uint8, device=self.device)
        self.step(noop)
        return self

    def turn(self):
        """env.py:65-66: len(moves) per env (uint8[N])."""
        out = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_observe(self.state.data_ptr(), None, None, out.data_ptr(), None, None,
                                             None, None, None, None, None, self.num_envs,
                                             _stream_ptr(self.device)))
        return out

    def observ(self):
        return self.observation()

    def observation(self, extras: bool = False, out=None):
        """env.py:68-85 as tensors: ``classical`` int8[N,9], ``q_states_p1`` int8[N,5,2],
        ``q_states_p2`` int8[N,4,2] (padded with -1), ``turn`` uint8[N]; with ``extras`` also
        ``moves`` int8[N,9,2], ``n_moves``, ``rounds`` (check_win), ``reward_p1`` (Env._reward),
        ``winner`` (0 none, 1 X, 2 O) and ``action_mask`` bool[N,36]."""
        return observe_states(self.state, extras=extras, out=out)

    def winner(self):
        """mcts.py:52-65 / strat_eval.py:21-32 per env: uint8[N], 0 none or draw, 1 X, 2 O."""
        out = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_observe(self.state.data_ptr(), None, None, None, None, None, None,
                                             None, None, out.data_ptr(), None, self.num_envs,
                                             _stream_ptr(self.device)))
        return out

    def action_mask(self):
        """mcts.py:87-91 for every env: bool[N,36] (``qttt_observe``'s ``mask_bool`` output)."""
        out = torch.empty((self.num_envs, 36), dtype=torch.bool, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_observe(self.state.data_ptr(), None, None, None, None, None, None,
                                             None, None, None, out.data_ptr(), self.num_envs,
                                             _stream_ptr(self.device)))
        return out

    def reward_p1(self):
        """``Env._reward()`` (env.py:87-112) per env: +1 X has the earlier line, -1 O, else 0."""
        out = torch.empty(self.num_envs, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.qttt_observe(self.state.data_ptr(), None, None, None, None, None, None,
                                             None, out.data_ptr(), None, None, self.num_envs,
                                             _stream_ptr(self.device)))
        return out

    # ------------------------------------------------------------------ helpers
    def load_positions(self, classical, moves, n_moves):
        """Sets every env to a reference-shaped position (Board.board, Board.moves)."""
        self.state.copy_(pack_states(classical, moves, n_moves, self.device))
        return self

    def _check_actions(self, actions):
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(actions, device=self.device)
        if actions.device != self.device:
            raise ValueError("actions must live on the env's device")
        if actions.dim() == 1:
            if actions.dtype != torch.uint8:
                # anything outside 0..35 is an illegal action; keep it one after the cast too
                # (260 must not wrap into the legal index 4)
                actions = torch.where((actions >= 0) & (actions < 36), actions,
                                      torch.full_like(actions, 255)).to(torch.uint8)
            fmt = _lib.ACT_INDEX
        elif actions.dim() == 2 and actions.shape[1] == 2:
            if actions.dtype != torch.int8:
                # values outside int8 would wrap; anything outside 0..8 is illegal anyway
                actions = actions.clamp(-1, 127).to(torch.int8)
            fmt = _lib.ACT_PAIR
        else:
            raise ValueError("actions must be uint8[N] indices or int8[N,2] pairs")
        if actions.shape[0] != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {actions.shape[0]}")
        return actions.contiguous(), fmt

    def render(self, index: int = 0) -> str:
        """``displayBoard`` text (display.py:4-32) of env ``index`` (env.py:59-60 for one env of the batch)."""
        return render_states(self.state, [int(index)])[0]

    def _obs(self):
        if self.obs_mode == "packed":
            return {"packed": self.state}
        return self.observation()

    def invalid(self):
        """bool[N]: the last ``step`` was a swallowed illegal action for this env (env.py:41-43)."""
        return (self.status & 3) == 1

    def _result(self):
        # no extra kernels here: everything returned is a buffer the step kernel wrote; the
        # derived entries of ``info`` are computed when (and only when) they are looked up
        info = _Info(self, {"action_mask": self.mask, "status": self.status})
        return self._obs(), self.reward, self.done, self._never, info


class _Info(dict):
    """``info`` of ``BatchedEnv.reset`` / ``step``.  ``action_mask`` (int64[N], bit k = action k
    is legal) and ``status`` are buffers the step kernel wrote.  The other entries SURVEY (b)
    lists are derived from the state on first lookup (one ``qttt_observe`` launch each), so a
    training loop that never reads them pays nothing:

    ``invalid``          bool[N]   the action was a swallowed illegal no-op (env.py:41-43)
    ``reward_p1``        f32[N]    ``Env._reward()`` (env.py:87-112)
    ``winner``           uint8[N]  0 none / draw, 1 X, 2 O (mcts.py:52-65)
    ``action_mask_bool`` bool[N,36] ``GameState.action_mask()`` (mcts.py:87-91)
    ``reset``            bool[N]   the env was auto-reset at the start of this step
    """
    _LAZY = ("invalid", "reward_p1", "winner", "action_mask_bool", "reset")

    def __init__(self, env, eager):
        super().__init__(eager)
        self._env = env

    def __missing__(self, key):
        env = self._env
        if key == "invalid":
            v = (env.status & 3) == 1
        elif key == "reset":
            v = (env.status & 4) != 0
        elif key == "reward_p1":
            v = env.reward_p1()
        elif key == "winner":
            v = env.winner()
        elif key == "action_mask_bool":
            v = env.action_mask()
        else:
            raise KeyError(key)
        self[key] = v
        return v

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._LAZY

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


def pack_actions(actions, choices):
    """uint8 action indices (0..35; anything else becomes an illegal action) and coin bits ->
    the one-byte-per-env input of ``step_host_packed``."""
    a = torch.where(actions < 36, actions, torch.full_like(actions, 63))
    return (a | ((choices & 1) << 7)).to(torch.uint8)


_LEGAL_OF_FREE = None


def _legal_table(device):
    """512-entry table: free-square set -> 36-bit legal mask (mcts.py:19-27)."""
    global _LEGAL_OF_FREE
    if _LEGAL_OF_FREE is None:
        tbl = []
        for m in range(512):
            v = 0
            for k, (i, j) in enumerate(PAIRS):
                if (m >> i) & 1 and (m >> j) & 1:
                    v |= 1 << k
            tbl.append(v)
        _LEGAL_OF_FREE = torch.tensor(tbl, dtype=torch.int64)
    return _LEGAL_OF_FREE.to(device)


def unpack_result(result):
    """int16 result words of ``step_host_packed`` -> (reward f32, terminated bool, mask int64,
    status uint8), identical to what ``step`` returns."""
    r = result.to(torch.int32) & 0xFFFF
    mask = _legal_table(result.device)[(r & 0x1FF).long()]
    terminated = ((r >> 9) & 1).bool()
    win = ((r >> 10) & 1).bool()
    neg_one = torch.tensor(-1.0, dtype=torch.float32, device=result.device)
    neg_zero = torch.tensor(-0.0, dtype=torch.float32, device=result.device)
    reward = torch.where(win, neg_one, neg_zero)
    status = ((r >> 11) & 3).to(torch.uint8)
    return reward, terminated, mask, status


def unpack_obs12(obs12):
    """The 12-byte records of ``step_host_obs12`` (int32[N,3]) -> ``(obs, reward, terminated, mask,
    status)``: ``obs`` = ``classical`` int8[N,9], ``q_states_p1`` int8[N,5,2], ``q_states_p2``
    int8[N,4,2] (padded with -1), ``turn`` uint8[N] as ``observation()`` returns them; the rest as
    ``unpack_result``.  Runs wherever ``obs12`` lives (the host, normally)."""
    dev = obs12.device
    w = obs12.to(torch.int64) & 0xFFFFFFFF
    w0, w1, w2 = w[:, 0], w[:, 1], w[:, 2]
    n = w.shape[0]
    nib = torch.stack([(w0 >> (4 * k)) & 15 for k in range(8)] + [w1 & 15], dim=1)          # value = owner + 1
    classical = (nib - 1).to(torch.int8)
    codes = torch.stack([(w1 >> (4 + 6 * k)) & 63 for k in range(4)] + [(w2 >> (6 * k)) & 63 for k in range(5)], dim=1)
    pairs = torch.full((64, 2), -1, dtype=torch.int8, device=dev)
    pairs[:36] = torch.tensor(PAIRS, dtype=torch.int8, device=dev)

    def qlist(slots, width):
        c = codes[:, slots]
        live = c < 36
        order = torch.sort((~live).to(torch.int8), dim=1, stable=True).indices     # live entries first, in slot order
        c = torch.where(live, c, torch.full_like(c, 63)).gather(1, order)
        return pairs[c][:, :width]
    q1, q2 = qlist([0, 2, 4, 6, 8], 5), qlist([1, 3, 5, 7], 4)
    n_moves = (nib != 0).sum(1) + (codes < 36).sum(1)
    obs = {"classical": classical, "q_states_p1": q1.contiguous(), "q_states_p2": q2.contiguous(),
           "turn": (n_moves & 1).to(torch.uint8)}
    free = ((nib == 0).to(torch.int64) << torch.arange(9, device=dev)).sum(1)
    mask = _legal_table(dev)[free]
    terminated = ((w1 >> 28) & 1).bool()
    win = ((w1 >> 29) & 1).bool()
    reward = torch.where(win, torch.tensor(-1.0, device=dev), torch.tensor(-0.0, device=dev))
    status = ((w1 >> 30) & 1).to(torch.uint8)
    return obs, reward, terminated, mask, status


def unpack_result12(result12, n: int):
    """The bit-packed result words of ``step_host_packed12`` (int16[3 * ceil(n / 4)]) -> the int16[n]
    words ``unpack_result`` takes: word k of a group holds env 4g+k's 12 bits, the three top nibbles
    together env 4g+3's."""
    w = (result12.to(torch.int32) & 0xFFFF).view(-1, 3)
    low = w & 0xFFF
    last = (w[:, 0] >> 12) | ((w[:, 1] >> 12) << 4) | ((w[:, 2] >> 12) << 8)
    out = torch.cat([low, last[:, None]], dim=1).reshape(-1)[:n]
    return out.to(torch.int16)


def observe_states(state, extras: bool = False, out=None):
    """``out``: the dict returned by an earlier call with the same shapes and ``extras`` -- its
    tensors are overwritten instead of allocating new ones."""
    lib = _lib.lib()
    n, dev = state.shape[0], state.device
    if out is not None:
        mask_u8 = out["action_mask"].view(torch.uint8) if "action_mask" in out else None
        with torch.cuda.device(dev):
            _lib.check(lib.qttt_observe(
                state.data_ptr(), out["classical"].data_ptr(), _lib.ptr(out.get("moves")),
                _lib.ptr(out.get("n_moves")), out["q_states_p1"].data_ptr(), out["q_states_p2"].data_ptr(),
                out["turn"].data_ptr(), _lib.ptr(out.get("rounds")), _lib.ptr(out.get("reward_p1")),
                _lib.ptr(out.get("winner")), _lib.ptr(mask_u8), n, _stream_ptr(dev)))
        return out
    e8 = lambda *shape: torch.empty(shape, dtype=torch.int8, device=dev)   # noqa: E731
    out = {"classical": e8(n, 9), "q_states_p1": e8(n, 5, 2), "q_states_p2": e8(n, 4, 2),
           "turn": torch.empty(n, dtype=torch.uint8, device=dev)}
    ex = {}
    if extras:
        ex = {"moves": e8(n, 9, 2), "n_moves": torch.empty(n, dtype=torch.uint8, device=dev),
              "rounds": e8(n, 2), "reward_p1": torch.empty(n, dtype=torch.float32, device=dev),
              "winner": torch.empty(n, dtype=torch.uint8, device=dev),
              "action_mask": torch.empty((n, 36), dtype=torch.uint8, device=dev)}
    with torch.cuda.device(dev):
        _lib.check(lib.qttt_observe(
            state.data_ptr(), out["classical"].data_ptr(), _lib.ptr(ex.get("moves")),
            _lib.ptr(ex.get("n_moves")), out["q_states_p1"].data_ptr(),
            out["q_states_p2"].data_ptr(), out["turn"].data_ptr(), _lib.ptr(ex.get("rounds")),
            _lib.ptr(ex.get("reward_p1")), _lib.ptr(ex.get("winner")),
            _lib.ptr(ex.get("action_mask")), n, _stream_ptr(dev)))
    if extras:
        ex["action_mask"] = ex["action_mask"].bool()
        out.update(ex)
    return out


def to_vector(state):
    """``GameState.to_vector`` (mcts.py:67-85) for packed states int32[N,4] -> float32[N,18,10]:
    the feature matrix the reference's policy/value net consumes (nn.py:30-42).  The reference
    computes it in float64; 1/sqrt(9) is rounded to float32 here (|diff| < 2e-8)."""
    lib = _lib.lib()
    n, dev = state.shape[0], state.device
    out = torch.empty((n, 18, 10), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.qttt_features(state.data_ptr(), out.data_ptr(), n, _stream_ptr(dev)))
    return out


def get_mask(state):
    """``nn.Model.get_mask`` (nn.py:44-61) for packed states int32[N,4] -> bool[N,36]: True where
    the action touches a classical square (the logits the reference's policy head sets to -inf);
    the complement of ``GameState.action_mask()`` (mcts.py:87-91)."""
    lib = _lib.lib()
    n, dev = state.shape[0], state.device
    out = torch.empty((n, 36), dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.qttt_get_mask(state.data_ptr(), out.data_ptr(), n, _stream_ptr(dev)))
    return out


def render_text(classical, moves, n_moves) -> str:
    """The 3x3-of-3x3 ASCII board of ``displayBoard`` (qtttgym/display.py:4-32) for ONE
    position given as reference-shaped lists: spooky mark of move i in sub-cell i of both its
    squares; a classical square shows its x / o pattern with the move index in the centre."""
    cells = [[" "] * 9 for _ in range(9)]
    for i in range(int(n_moves)):
        a, b = int(moves[i][0]), int(moves[i][1])
        cells[a][i] = str(i)
        cells[b][i] = str(i)
    for sq, owner in enumerate(classical):
        owner = int(owner)
        if owner >= 0:
            mark = "x" if owner % 2 == 0 else "o"
            for j in range(9):
                cells[sq][j] = mark if j % 2 == owner % 2 else " "
            cells[sq][4] = str(owner)
    out = ""
    for i in range(3):
        out += "+---+---+---+\n"
        for k in range(3):
            for j in range(3):
                out += "|" + "".join(cells[3 * i + j][3 * k:3 * k + 3])
            out += "|\n"
    return out + "+---+---+---+\n"


def render_states(state, indices=None):
    """ASCII dumps (``displayBoard`` layout, display.py:4-32) of packed states int32[N,4]: a list
    with one string per requested game (all of them when ``indices`` is None) -- the debugging aid
    for parity failures.  One ``qttt_observe`` launch for the selection."""
    if indices is not None:
        state = state[torch.as_tensor(indices, dtype=torch.long, device=state.device)].contiguous()
    if state.shape[0] == 0:
        return []
    obs = observe_states(state, extras=True)
    classical, moves, n_moves = (obs[k].cpu().tolist() for k in ("classical", "moves", "n_moves"))
    out = []
    for cl, mv, nm in zip(classical, moves, n_moves):
        # the autofill entry is stored as (s, s) like the reference's (s, s, 8) (board.py:25)
        out.append(render_text(cl, mv, nm))
    return out


def pack_states(classical, moves, n_moves, device="cuda"):
    """Reference-shaped positions -> packed states int32[N,4] (see qttt_pack)."""
    lib = _lib.lib()
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    classical = torch.as_tensor(classical, dtype=torch.int8).to(dev).contiguous()
    moves = torch.as_tensor(moves, dtype=torch.int8).to(dev).contiguous()
    n_moves = torch.as_tensor(n_moves, dtype=torch.uint8).to(dev).contiguous()
    n = classical.shape[0]
    assert classical.shape == (n, 9) and moves.shape == (n, 9, 2) and n_moves.shape == (n,)
    state = torch.empty((n, 4), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.qttt_pack(state.data_ptr(), classical.data_ptr(), moves.data_ptr(),
                                 n_moves.data_ptr(), n, _stream_ptr(dev)))
    return state


class Env:
    """Single-env adapter with the reference's exact return types (qtttgym/env.py:15-66):
    Python lists / tuples in ``obs``, a Python float reward (``-0.0`` / ``-1.0``), bools, and the
    reference's ``action_space`` / ``observation_space`` attributes.

    One ``step`` is ONE kernel launch (``qttt_env1``): the action travels as kernel arguments, the
    kernel steps the game, decodes everything ``step`` / ``observ`` / ``turn`` / ``_reward`` report
    into a 128-byte record and writes it straight into mapped pinned host memory; the host spins
    on the record's sequence word.  No copy engine, no stream synchronisation: the latency of a
    step is launch + ~one PCIe write.  Throughput comes from ``BatchedEnv``.

    Deviations, all documented in DESIGN.md: ``obs["classical"]`` is a snapshot list, not an
    alias of live state (Q5); the collapse coin comes from Philox ``(seed, 0, len(moves), episode)``
    unless ``coin`` is passed to ``step`` (``seed=None`` draws the seed from the OS like the
    reference's unseeded ``random``); the classical range of ``observation_space`` is corrected
    to -1..8 (Q6)."""

    # byte offsets inside the record (include/qttt_b200.h: qttt_env1)
    _STATE, _MASK, _REWARD, _DONE, _STATUS, _TURN, _NMOVES = 0, 16, 24, 28, 29, 30, 31
    _CLASSICAL, _Q1, _Q2, _ROUNDS, _WINNER, _REWARD_P1, _MOVES, _SEQ = 32, 48, 58, 66, 68, 72, 80, 124
    _BYTES = 128

    def __init__(self, device="cuda", seed=None):
        import random as _random
        import struct
        from . import spaces as _spaces
        self.lib = _lib.lib()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("qtttgym_b200 runs on CUDA devices only (no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self._device = dev
        self.seed = (_random.SystemRandom().getrandbits(64) if seed is None else int(seed)) & 0xFFFFFFFFFFFFFFFF
        self.epoch = 0
        self.action_space = _spaces.action_space()              # env.py:19
        self.observation_space = _spaces.observation_space()    # env.py:20-25 (Q6 corrected)
        self._state = torch.zeros((1, 4), dtype=torch.int32, device=dev)
        self._host = torch.zeros(self._BYTES, dtype=torch.uint8).pin_memory()
        self._np = self._host.numpy()
        self._seq_view = self._np[self._SEQ:self._SEQ + 4].view("uint32")
        self._mv = memoryview(self._np).cast("b")
        self._unpack_f = struct.Struct("<f").unpack_from
        self._seq = 0
        self._last_status = 0
        self._first = True
        # everything a call needs, resolved once: a step is latency-bound, and looking these up
        # through torch on every call costs more than the kernel does
        self._fn = self.lib.qttt_env1
        self._state_ptr = self._state.data_ptr()
        self._host_ptr = self._host.data_ptr()
        self._dev_index = dev.index
        self._stream = _stream_ptr(dev)          # the stream that is current now (the default stream)
        self.reset()

    # -- plumbing ---------------------------------------------------------------------------
    def _call(self, op, a=0, b=0, coin=-1):
        seq = self._seq = (self._seq % 0xFFFFFFFE) + 1
        if torch.cuda.current_device() == self._dev_index:
            rc = self._fn(self._state_ptr, op, a, b, coin, self.seed, self.epoch, self._host_ptr, seq, self._stream)
        else:
            with torch.cuda.device(self._device):
                rc = self._fn(self._state_ptr, op, a, b, coin, self.seed, self.epoch, self._host_ptr, seq,
                              self._stream)
        if rc:
            _lib.check(rc)
        _lib.LAUNCHES += 1
        view = self._seq_view
        spins = 0
        while view[0] != seq:                    # the kernel writes the sequence word last
            spins += 1
            if spins > 5_000_000:                # ~seconds: the launch failed asynchronously
                torch.cuda.synchronize(self._device)
                if view[0] != seq:
                    raise RuntimeError("qttt_env1: the record never arrived")

    def _observation(self):
        mv = self._mv
        q1, q2, cl = self._Q1, self._Q2, self._CLASSICAL
        return {"q_states_p1": [(mv[q1 + 2 * k], mv[q1 + 2 * k + 1]) for k in range(5) if mv[q1 + 2 * k] >= 0],
                "q_states_p2": [(mv[q2 + 2 * k], mv[q2 + 2 * k + 1]) for k in range(4) if mv[q2 + 2 * k] >= 0],
                "classical": list(mv[cl:cl + 9]),
                "turn": mv[self._TURN]}

    # -- the reference API ------------------------------------------------------------------
    def reset(self, *, seed=None, options=None):
        """env.py:55-57 (seed / options ignored, Q4).  A new episode: the Philox epoch advances, so
        its collapses are independent of the previous episode's."""
        if self._first:
            self._first = False
        else:
            self.epoch += 1
        self._call(1)
        return self._observation(), {}

    def step(self, action, verbose=False, coin=None):
        """env.py:34-53."""
        try:
            a, b = int(action[0]), int(action[1])
            if not (0 <= a < 9 and 0 <= b < 9):
                a = b = -1                  # IndexError path of the reference: a swallowed no-op
        except Exception as e:              # env.py:41: anything raised becomes a no-op
            if verbose:
                print("noop (i.e. invalid) move...", e)
            a = b = -1
        self._call(0, a, b, -1 if coin is None else int(coin) & 1)
        mv = self._mv
        self._last_status = mv[self._STATUS]
        if verbose and self._last_status == 1:
            print("noop (i.e. invalid) move...")
        return self._observation(), self._unpack_f(self._np, self._REWARD)[0], bool(mv[self._DONE]), False, {}

    def observ(self):
        return self._observation()

    def turn(self):
        """env.py:65-66."""
        return self._mv[self._NMOVES]

    def action_mask(self):
        """mcts.py:87-91: bool[36]."""
        m = int(self._np[self._MASK:self._MASK + 8].view("uint64")[0])
        import numpy as np
        return np.array([(m >> k) & 1 for k in range(36)], dtype=bool)

    def render(self):
        """env.py:59-60 -> displayBoard (display.py:4-32)."""
        i8 = self._np.view("int8")
        print(render_text(i8[self._CLASSICAL:self._CLASSICAL + 9].tolist(),
                          i8[self._MOVES:self._MOVES + 18].reshape(9, 2).tolist(), self.turn()))

    def _reward(self):
        """env.py:87-112."""
        return self._unpack_f(self._np, self._REWARD_P1)[0]

    def winner(self):
        """0 none / draw, 1 X, 2 O (mcts.py:52-65)."""
        return self._mv[self._WINNER]

    @property
    def state(self):
        """the packed state as an int32[1,4] device tensor"""
        return self._state


__all__ = ["BatchedEnv", "Env", "observe_states", "pack_states", "pack_actions", "unpack_result",
           "to_vector", "get_mask", "render_text", "PAIRS"]
