"""Times (and, under ncu, drives) the tree search over 32,768 roots: 100 rollouts x 10 playouts.

    python profiles/mcts_drive.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    import torch
    import qtttgym_b200 as Q
    env = Q.BatchedEnv(32768, seed=20261018)
    for _ in range(4):
        env.step_random()
    roots = env.state.clone()
    mc = Q.BatchedMCTS(rollouts=100, num_simulations=10, seed=1)

    def run():
        mc.reset(roots, total_rollouts=100)
        mc.contemplate(100)
    run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"32768 roots x 100 rollouts x 10 sims: {ms:.2f} ms, {32768 * 100 / ms / 1e3:.1f} M rollouts/s", flush=True)


if __name__ == "__main__":
    main()
