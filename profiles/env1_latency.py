"""Where the time of one qtttgym_b200.Env.step goes (run on the GPU box)."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import qtttgym_b200 as Q

env = Q.Env(seed=1)
N = 20000
def t(fn, n=N):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6
print("call+spin (op 2: re-emit record)   %.2f us" % t(lambda: env._call(2)))
print("call+spin (op 0: illegal no-op step) %.2f us" % t(lambda: env._call(0, -1, -1, 0)))
print("_observation() decode               %.2f us" % t(env._observation))
print("step((0,0)) full                    %.2f us" % t(lambda: env.step((0, 0))))
board = [-1] * 9
rng = random.Random(1)
def legal_choice():
    legal = [p for p in Q.PAIRS if board[p[0]] == -1 and board[p[1]] == -1]
    return rng.choice(legal)
print("user loop part (legal list + choice) %.2f us" % t(legal_choice))
# the reference-style game loop
t0 = time.perf_counter(); n = 0
for _ in range(500):
    obs, _ = env.reset(); term = False
    while not term:
        b = obs["classical"]
        legal = [p for p in Q.PAIRS if b[p[0]] == -1 and b[p[1]] == -1]
        obs, r, term, _, _ = env.step(rng.choice(legal)); n += 1
print("game loop: %.2f us per step (%d steps)" % ((time.perf_counter() - t0) / n * 1e6, n))
# kernel duration via events
a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(1000):
    env._fn(env._state_ptr, 2, 0, 0, -1, env.seed, env.epoch, env._host_ptr, 1, env._stream)
b2.record(); torch.cuda.synchronize()
print("k_env1 back-to-back on the device: %.2f us per launch" % (a.elapsed_time(b2)))
