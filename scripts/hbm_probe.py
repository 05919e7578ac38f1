"""HBM bandwidth by access mix: write-only (fill), read-only (sum), copy.  The step kernel's
DRAM traffic is ~76 % writes, so the write-only figure is the one that bounds it."""
import torch
dev = torch.device("cuda:0")
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
def timeit(fn, nbytes, reps=10):
    for _ in range(3): fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return nbytes / best / 1e6
print("fill  (write only) GB/s: %.0f" % timeit(lambda: a.fill_(1), n))
print("zero  (memset)     GB/s: %.0f" % timeit(lambda: a.zero_(), n))
a32 = a.view(torch.int32)
print("sum   (read only)  GB/s: %.0f" % timeit(lambda: a32.sum(), n))
print("copy  (r+w)        GB/s: %.0f" % timeit(lambda: b.copy_(a), 2 * n))
# 1 read : 3 write mix: expand a quarter-size input into the full output (repeat_interleave-like)
q = a[: n // 4].view(n // 4, 1)
out = b.view(n // 4, 4)
print("1r:3w expand       GB/s: %.0f" % timeit(lambda: torch.add(q, 0, out=out) if False else out.copy_(q.expand(n // 4, 4)), n + n // 4))
