"""Why does the step kernel take longer in the loop than under ncu?  Per-kernel event times
at different points of a run, with and without synchronisation between launches."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
n = 1 << 20
dev = torch.device("cuda:0")
b = ph.HanabiBatch(dict(players=2), n, (np.arange(n) + 1).astype(np.int32))
obs = torch.empty((n, b.obs_size), dtype=torch.uint8, device=dev)
mask = torch.empty((n, b.max_moves), dtype=torch.uint8, device=dev)
cur = torch.empty(n, dtype=torch.int32, device=dev); rew = torch.empty(n, dtype=torch.int32, device=dev)
done = torch.empty(n, dtype=torch.uint8, device=dev); act = torch.empty(n, dtype=torch.int32, device=dev)
b.reset(); b.observe(obs, mask, cur)
select = ph.bind_select_action(None, mask, 1.0, seed=7, out=act)
step = b.bind_step_encode(act, rew, done, obs, mask, cur)
def run(k, base, sync):
    evs = []
    for t in range(k):
        select(base + t)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        if sync: torch.cuda.synchronize()
        e0.record(); step(); e1.record()
        if sync: torch.cuda.synchronize()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(c) for a, c in evs]) * 1e3
t = 0
for label, k, sync in [("first 20 (sync)", 20, True), ("next 20", 20, False), ("next 300", 300, False), ("next 20 (sync)", 20, True),
                       ("next 1000", 1000, False), ("next 20 (sync)", 20, True)]:
    us = run(k, t, sync); t += k
    st = b.read_stats()
    print(f"{label:18s} step-kernel us: mean {us.mean():7.1f} min {us.min():7.1f} max {us.max():7.1f}  first5 {np.round(us[:5],0)}  done/step {st['episodes']/max(t,1)/n:.3f}")
# L2-clean experiment: flush L2 by writing a big buffer between kernels
junk = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
us = []
for i in range(10):
    select(t + i); junk.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize(); us.append(e0.elapsed_time(e1) * 1e3)
print("after 512MB memset (L2 full of other dirty lines):", np.round(us, 0))

import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def nv():
    return (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
            pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetTemperature(h, 0),
            hex(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
print("idle", nv())
torch.cuda.synchronize(); time.sleep(2.0)
print("after 2s idle", nv())
for seg in range(12):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(250):
        select(t); step(); t += 1
    e1.record()
    mid = nv()
    torch.cuda.synchronize()
    print(f"seg {seg}: {e0.elapsed_time(e1)/250*1e3:7.1f} us/step  nvml(sm,mem,W,C,reasons)={mid}")
