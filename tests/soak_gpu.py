"""One-off GPU soak (not collected by pytest): `PYTHONPATH=. python tests/soak_gpu.py` on a B200.
16 x 2^20 games through the step API in both action formats, with and without illegal actions
and post-terminal moves, every output after every ply against the C oracle; then 2^22 qeval
boards, a 2e7-game sweep and 4096 x 512 rollouts; round 2 adds 2^20-env auto-resetting batches
in both modes, the fused step + observation launch and the 12-bit mapped results.
Round 1: 140,646,913 accepted steps, 0 mismatches, 167 s."""
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import parity_suite as S
from backends import CudaBackend
cuda = CudaBackend()
t0 = time.time(); total = 0
for b in range(16):
    total += S.check_random_play(cuda, 1 << 20, 5000 + b, illegal_rate=(0.0, 0.03, 0.1, 0.0)[b % 4],
                                 overrun=(b % 4 == 2), fmt=("pair", "index")[b % 2], full_obs_every=3)
    print(b, total, round(time.time() - t0, 1), flush=True)
for mode in ("apply", "next"):                     # desynchronised auto-resetting batches
    S.check_autoreset(cuda, 1 << 20, seed=7, mode=mode, steps=30)
S.check_step_obs(cuda, 1 << 20, seed=9)            # the fused step + observation launch, all modes
S.check_packed_step(cuda, (1 << 20) + 3, seed=11, variant="mapped12")   # 12-bit results over mapped memory
print("autoreset / step_obs / packed12 ok", round(time.time() - t0, 1), flush=True)
S.check_qeval_both(cuda, 1 << 22, 99)
S.check_sweep(cuda, 20_000_000, 77)
S.check_rollout(cuda, 4096, 512, 5)
print("soak ok", total, "accepted steps", round(time.time() - t0, 1), "s")
