import sys, os
sys.path.insert(0, os.getcwd())
import torch, qtttgym_b200 as Q
dev = torch.device("cuda", 0)
for _ in range(2): Q.selfplay_sweep(0, 125_000_000, 1, device=dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(5): st = Q.selfplay_sweep(0, 125_000_000, 1, device=dev)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(os.environ.get("QTTT_B200_LIB", "default")[-14:], "sweep ms", round(ms, 3), "G steps/s", round(int(st[3]) / ms / 1e6, 1))
roots = torch.zeros((65536, 4), dtype=torch.int32, device=dev)
o = Q.rollout_eval(roots, 256, 1)
torch.cuda.synchronize(); a.record()
for _ in range(5): Q.rollout_eval(roots, 256, 1, out=o)
b.record(); torch.cuda.synchronize()
print("rollout from empty roots 65536x256 ms", round(a.elapsed_time(b) / 5, 3))
