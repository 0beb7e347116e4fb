"""Scratch: write-only HBM bandwidth on this GPU (torch fill / zero_) next to the copy bandwidth the roofline uses."""
import torch
dev = "cuda:0"
n = 14 * 10**9 // 4
x = torch.empty(n, dtype=torch.float32, device=dev)
y = torch.empty(n // 2, dtype=torch.float32, device=dev)
z = torch.empty(n // 2, dtype=torch.float32, device=dev)
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1.5)); print("fill_ 14 GB: %.3f ms  %.0f GB/s written" % (ms, x.numel() * 4 / ms / 1e6))
ms = t(lambda: x.zero_()); print("zero_ 14 GB: %.3f ms  %.0f GB/s written" % (ms, x.numel() * 4 / ms / 1e6))
ms = t(lambda: y.copy_(z)); print("copy 7+7 GB: %.3f ms  %.0f GB/s read+write" % (ms, 2 * y.numel() * 4 / ms / 1e6))
ms = t(lambda: torch.sum(x)); print("sum 14 GB: %.3f ms  %.0f GB/s read" % (ms, x.numel() * 4 / ms / 1e6))
