"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute)."""
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"QTTT_API[^;(]*?\b(qttt_\w+)\s*\(", text))
    return names


@pytest.fixture(scope="module")
def built_lib():
    from qtttgym_b200 import build
    return build.build()


def test_header_declares_the_expected_surface():
    assert _declared_symbols() == {
        "qttt_abi_version", "qttt_strerror", "qttt_reset", "qttt_reset_all", "qttt_reset_step", "qttt_step", "qttt_step_ex",
        "qttt_step_packed", "qttt_step_packed_obs", "qttt_step_packed_mapped", "qttt_step_packed12_mapped", "qttt_step_packed12_host", "qttt_step_packed_host_obs12", "qttt_step_packed_host",
        "qttt_step_packed_host_obs", "qttt_step_random", "qttt_step_random_ex",
        "qttt_observe", "qttt_features", "qttt_env1", "qttt_qeval1", "qttt_get_mask", "qttt_step_features", "qttt_step_obs", "qttt_pack", "qttt_qeval_both", "qttt_rollout", "qttt_sweep",
        "qttt_mcts_node_bytes", "qttt_mcts_init", "qttt_mcts_run", "qttt_mcts_stats", "qttt_mcts_sync"}


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert _declared_symbols() <= exported
    assert not any(s.startswith("orc_") or s.startswith("emu_") for s in exported)   # no oracle / emulation inside


def test_ctypes_binding_matches_header(built_lib):
    from qtttgym_b200 import _lib
    assert set(_lib.EXPORTED_SYMBOLS) == _declared_symbols()
    lib = _lib.lib()
    assert lib.qttt_abi_version() == 2
    assert lib.qttt_strerror(0) == b"ok"
    assert b"invalid argument" in lib.qttt_strerror(-1)
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.qttt_reset(None, None, 4, None) == -1
    assert lib.qttt_sweep(3, 2, 0, None, None) == -1


def test_sass_is_sm100_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_in_package():
    """The product never imports the oracle or the host emulation (docstrings may cite them)."""
    import ast
    pkg = os.path.join(ROOT, "qtttgym_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for name in names:
                assert not name.split(".")[0] in ("oracle", "hostemu", "tests"), (path, name)
        text = open(path).read()
        assert "libqttt_oracle" not in text and "libqttt_hostemu" not in text and "CDLL(" not in text.replace(
            "C.CDLL(LIB)", ""), path
    for path in glob.glob(os.path.join(pkg, "csrc", "*.cu*")):
        text = open(path).read()
        assert "qttt_oracle" not in text and "hostemu.cpp" not in text, path


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from qtttgym_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.QtttLibraryError):
        _lib.lib()


def test_action_table():
    import qtttgym_b200 as Q
    assert len(Q.PAIRS) == 36 and Q.PAIRS[0] == (0, 1) and Q.PAIRS[7] == (0, 8) and Q.PAIRS[35] == (7, 8)
    for k, (i, j) in enumerate(Q.PAIRS):
        assert Q.ind2move(k) == (i, j) and Q.move2ind(i, j) == k == Q.move2ind(j, i)


def test_cpu_device_is_refused():
    import qtttgym_b200 as Q
    with pytest.raises(RuntimeError):
        Q.BatchedEnv(4, device="cpu")


def test_spaces_mirror_the_reference_env():
    """qtttgym/env.py:19-25: action_space = Tuple(Discrete(9), Discrete(9)); the observation Dict
    (classical range corrected to -1..8, quirk Q6)."""
    import numpy as np
    from qtttgym_b200 import spaces as S
    a = S.action_space()
    assert len(a) == 2 and a[0].n == 9 and a[1].n == 9
    for _ in range(50):
        x = a.sample()
        assert a.contains(x) and x in a
    assert not a.contains((9, 0)) and not a.contains((0,))
    o = S.observation_space()
    obs = {"q_states_p1": [(0, 1)], "q_states_p2": [], "classical": np.array([-1, 8, 0, 1, 2, 3, 4, 5, 6], np.int32),
           "turn": 1}
    if not S.HAVE_GYMNASIUM:
        assert o.contains(obs)
        assert o["q_states_p1"].max_len == 5 and o["q_states_p2"].max_len == 4
        assert not o["classical"].contains(np.full(9, 9, np.int32))


def test_unpack_result12_is_the_inverse_of_the_kernel_packing():
    """host-side decode of the 12-bit result words (pure torch, no GPU)"""
    import numpy as np
    import torch
    from qtttgym_b200.env import unpack_result12
    rng = np.random.default_rng(0)
    for n in (1, 3, 4, 5, 1023, 4096):
        r = rng.integers(0, 1 << 12, n).astype(np.uint32)
        pad = np.concatenate([r, np.zeros((-n) % 4, np.uint32)]).reshape(-1, 4)
        w = np.stack([pad[:, 0] | ((pad[:, 3] & 15) << 12), pad[:, 1] | (((pad[:, 3] >> 4) & 15) << 12),
                      pad[:, 2] | (((pad[:, 3] >> 8) & 15) << 12)], 1).astype(np.uint16).reshape(-1)
        got = unpack_result12(torch.from_numpy(w.view(np.int16).copy()), n).numpy().view(np.uint16)
        assert np.array_equal(got.astype(np.uint32), r)


def test_unpack_obs12_decodes_a_hand_built_record():
    """host-side decode of the 12-byte observation records (pure torch, no GPU)"""
    import torch
    from qtttgym_b200.env import unpack_obs12

    def ind(a, b):
        return (15 * a - a * a + 2 * b - 2) // 2
    # squares 0, 1, 2 classical (owners 2, 0, 1); uncollapsed: move 3 = (3, 4) [O], move 4 = (5, 6) [X];
    # flags: illegal no-op
    codes = [63] * 9
    codes[3], codes[4] = ind(3, 4), ind(5, 6)
    w0 = 3 | (1 << 4) | (2 << 8)
    w1 = sum(codes[k] << (4 + 6 * k) for k in range(4)) | (1 << 30)
    w2 = sum(codes[4 + k] << (6 * k) for k in range(5))
    rec = torch.tensor([[w0, w1, w2]], dtype=torch.int64).to(torch.int32)
    obs, reward, term, mask, status = unpack_obs12(rec)
    assert obs["classical"].tolist() == [[2, 0, 1, -1, -1, -1, -1, -1, -1]]
    assert obs["q_states_p1"].tolist() == [[[5, 6]] + [[-1, -1]] * 4]
    assert obs["q_states_p2"].tolist() == [[[3, 4]] + [[-1, -1]] * 3]
    assert obs["turn"].tolist() == [1] and term.tolist() == [False] and status.tolist() == [1]
    assert reward.view(torch.int32).tolist() == [-(1 << 31)]          # -0.0
    assert bin(int(mask[0])).count("1") == 15                          # pairs of the 6 free squares
