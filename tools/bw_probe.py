"""Write-only / read-only / copy HBM bandwidth with plain torch kernels (roofline denominators for output-dominated ops)."""
import torch
dev = "cuda:0"
n = 1 << 29   # 512 Mi halves = 1 GiB
a = torch.empty(n, dtype=torch.float16, device=dev)
b = torch.empty(n, dtype=torch.float16, device=dev)
def timeit(f, reps=10):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
gb = n * 2 / 1e9
t = timeit(lambda: a.fill_(1.0)); print(f"write-only fill_: {gb / t * 1e3:.0f} GB/s")
t = timeit(lambda: a.zero_()); print(f"write-only zero_ (memset): {gb / t * 1e3:.0f} GB/s")
t = timeit(lambda: torch.sum(a)); print(f"read-only sum: {gb / t * 1e3:.0f} GB/s")
t = timeit(lambda: b.copy_(a)); print(f"copy (read+write bytes): {2 * gb / t * 1e3:.0f} GB/s")
t = timeit(lambda: torch.add(a, b, out=b)); print(f"add a+b->b (2 reads + 1 write): {3 * gb / t * 1e3:.0f} GB/s")
