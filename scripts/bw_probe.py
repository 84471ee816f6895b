"""GPU measurement aid: read-only vs copy HBM bandwidth with plain torch kernels (context for roofline.frac)."""
import torch
n = 4 * 1024 ** 3            # 4 Gi doubles = 32 GiB
x = torch.ones(n, dtype=torch.float64, device="cuda")
def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best
t = timed(lambda: torch.sum(x))
print(f"read-only  (torch.sum, 32 GiB f64): {n * 8 / t / 1e9:.0f} GB/s  ({t * 1e3:.2f} ms)")
y = torch.empty(n // 2, dtype=torch.float64, device="cuda")
t = timed(lambda: y.copy_(x[: n // 2]))
print(f"copy       (16 GiB -> 16 GiB):      {n * 8 / t / 1e9:.0f} GB/s read+write")
t = timed(lambda: y.fill_(2.0))
print(f"write-only (fill 16 GiB):           {n * 4 / t / 1e9:.0f} GB/s")
