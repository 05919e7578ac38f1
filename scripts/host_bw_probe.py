"""How much can this host write?  HostExpandPacked (CPU, non-temporal stores) alone, a pinned
D2H copy alone, and both at once -- to see whether DMA ingest and CPU stores share one limit."""
import json, sys, threading, time
import torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
assert ph.try_cdef() and ph.try_load()
n, dim = 1 << 20, 658
bits = torch.randint(-2**31, 2**31 - 1, (n, 24), dtype=torch.int32).pin_memory()
out = torch.empty((n, dim), dtype=torch.uint8).pin_memory()
d_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
h_buf = torch.empty(512 << 20, dtype=torch.uint8).pin_memory()
def expand(reps):
    t0 = time.perf_counter()
    for _ in range(reps): ph.host_expand_packed(bits, dim, out=out)
    return n * dim * reps / (time.perf_counter() - t0) / 1e9
def d2h(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): h_buf.copy_(d_buf, non_blocking=True)
    torch.cuda.synchronize()
    return d_buf.numel() * reps / (time.perf_counter() - t0) / 1e9
res = {"threads": ph.host_policy_threads()}
expand(2); d2h(2)
res["expand_alone_gbs"] = expand(10)
res["d2h_alone_gbs"] = d2h(10)
box = {}
def run_d2h(): box["d2h"] = d2h(14)
th = threading.Thread(target=run_d2h); th.start()
box["expand"] = expand(10); th.join()
res["expand_with_d2h_gbs"], res["d2h_with_expand_gbs"] = box["expand"], box["d2h"]
print(json.dumps(res))
