"""Device time of the C51 kernels alone: BatchSelectActionC51 on padded bf16 GEMM outputs
(2^18 games x 20 actions x 51 atoms, pitch 1024; u17 streaming kernel vs the row-staged one),
BatchProjectDistribution and BatchC51TargetDistribution at B = 32 and 32768."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
dev = torch.device("cuda:0")
def timed(fn, k=50, flush=None):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(k):
        if flush is not None: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / k * 1e3  # us
out = {}
n, A, atoms, pitch = 1 << 18, 20, 51, 1024
torch.manual_seed(0)
logits = (torch.randn((n, pitch), device=dev) * 2).to(torch.bfloat16)
mask = (torch.rand((n, A), device=dev) < 0.5).to(torch.uint8); mask[:, 3] = 1
support = torch.linspace(-25.0, 25.0, atoms, device=dev)
act = torch.empty(n, dtype=torch.int32, device=dev)
sel = ph.bind_select_action_c51(logits, atoms, mask, 0.0, seed=7, out=act, support=support)
us = timed(lambda i: sel(i))
bytes_in = n * (pitch * 2 + A) + n * 4
out["select_c51"] = {"kernel": os.environ.get("HANABI_B200_SELECT_U17", "1") != "0" and "k_select_c51_u17" or "k_select_c51_thirds",
                     "games": n, "us": us, "GB_per_s": bytes_in / us / 1e3,
                     "note": "2^18 x 1024 bf16 logits (537 MB) > L2, back to back"}
# reference check against fp32 math on a sample
idx = torch.arange(0, n, 997, device=dev)
lg = logits[idx][:, :A * atoms].float().view(-1, A, atoms)
q = (torch.softmax(lg, dim=2) * support).sum(dim=2)
q = torch.where(mask[idx] != 0, q, torch.full_like(q, -float("inf")))
want = q.argmax(dim=1).to(torch.int32)
sel(0); torch.cuda.synchronize()
top2 = q.topk(2, dim=1).values
clear = (top2[:, 0] - top2[:, 1]) > 1e-3
out["select_c51"]["argmax_agreement_clear_rows"] = float((act[idx][clear] == want[clear]).float().mean())
for B in (32, 32768):
    z = torch.linspace(-25, 25, 51, device=dev)
    s = torch.randn((B, 51), device=dev) * 12
    w = torch.softmax(torch.randn((B, 51), device=dev), dim=1)
    o = torch.empty((B, 51), device=dev)
    us = timed(lambda i: ph.project_distribution(s, w, z, out=o), k=100)
    out["project_B%d" % B] = {"us": us, "bytes": B * 51 * 4 * 3, "GB_per_s": B * 51 * 12 / us / 1e3}
    r = torch.randn(B, device=dev); t = (torch.rand(B, device=dev) < 0.2).to(torch.uint8)
    nl = torch.randn((B, 20, 51), device=dev) * 2
    nm = (torch.rand((B, 20), device=dev) < 0.5).to(torch.uint8); nm[:, 0] = 1
    us = timed(lambda i: ph.c51_target_distribution(r, t, nl, nm, z, 0.99), k=100)
    out["c51_target_B%d" % B] = {"us": us, "bytes": B * (20 * 51 * 4 + 51 * 4 + 29), "GB_per_s": B * (20 * 51 * 4 + 51 * 4 + 29) / us / 1e3}
print(json.dumps(out))
