import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = t(lambda: x.fill_(3)); print("fill 1 GiB: %.1f us  %.0f GB/s written" % (ms * 1e3, (1 << 30) / ms / 1e6))
ms = t(lambda: x.view(torch.int32).fill_(3)); print("fill int32: %.1f us  %.0f GB/s" % (ms * 1e3, (1 << 30) / ms / 1e6))
ms = t(lambda: torch.cuda.current_stream().synchronize() or x.zero_()); print("zero_: %.1f us %.0f GB/s" % (ms * 1e3, (1 << 30) / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy 1 GiB: %.1f us  %.0f GB/s (r+w)" % (ms * 1e3, 2 * (1 << 30) / ms / 1e6))
ms = t(lambda: x.view(torch.int32).sum()); print("read-sum: %.1f us %.0f GB/s" % (ms * 1e3, (1 << 30) / ms / 1e6))
