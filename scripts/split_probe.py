"""How does the fused step kernel's time decompose?  Times BatchStep (no outputs but reward/done),
BatchObserve (obs + mask + cur) and the fused BatchStepEncode on the same workload."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
n = 1 << 20
dev = torch.device("cuda:0")
def setup():
    b = ph.HanabiBatch(dict(players=2), n, (np.arange(n) + 1).astype(np.int32))
    t = dict(obs=torch.empty((n, b.obs_size), dtype=torch.uint8, device=dev), mask=torch.empty((n, b.max_moves), dtype=torch.uint8, device=dev),
             cur=torch.empty(n, dtype=torch.int32, device=dev), rew=torch.empty(n, dtype=torch.int32, device=dev),
             done=torch.empty(n, dtype=torch.uint8, device=dev), act=torch.empty(n, dtype=torch.int32, device=dev))
    b.reset(); b.observe(t["obs"], t["mask"], t["cur"])
    return b, t
def ev(): return torch.cuda.Event(True), torch.cuda.Event(True)
# split path
b, t = setup()
sel = ph.bind_select_action(None, t["mask"], 1.0, seed=7, out=t["act"])
ts, to = [], []
for i in range(400):
    sel(i)
    a0, a1 = ev(); a0.record(); b.step(t["act"], t["rew"], t["done"]); a1.record()
    c0, c1 = ev(); c0.record(); b.observe(t["obs"], t["mask"], t["cur"]); c1.record()
    if i >= 300: ts.append((a0, a1)); to.append((c0, c1))
torch.cuda.synchronize()
print("split:  BatchStep %.1f us   BatchObserve %.1f us" % (np.mean([x.elapsed_time(y) for x, y in ts]) * 1e3, np.mean([x.elapsed_time(y) for x, y in to]) * 1e3))
# fused
b, t = setup()
sel = ph.bind_select_action(None, t["mask"], 1.0, seed=7, out=t["act"])
f = b.bind_step_encode(t["act"], t["rew"], t["done"], t["obs"], t["mask"], t["cur"])
tf = []
for i in range(400):
    sel(i); a0, a1 = ev(); a0.record(); f(); a1.record()
    if i >= 300: tf.append((a0, a1))
torch.cuda.synchronize()
print("fused:  BatchStepEncode %.1f us" % (np.mean([x.elapsed_time(y) for x, y in tf]) * 1e3))
# pure write bandwidth reference: memset of the obs tensor
a0, a1 = ev(); a0.record()
for _ in range(20): t["obs"].fill_(1)
a1.record(); torch.cuda.synchronize()
print("fill 690 MB: %.1f us (%.2f TB/s)" % (a0.elapsed_time(a1) / 20 * 1e3, t["obs"].numel() / (a0.elapsed_time(a1) / 20 * 1e-3) / 1e12))
src = torch.empty_like(t["obs"])
a0, a1 = ev(); a0.record()
for _ in range(20): t["obs"].copy_(src)
a1.record(); torch.cuda.synchronize()
print("copy 690 MB: %.1f us (%.2f TB/s r+w)" % (a0.elapsed_time(a1) / 20 * 1e3, 2 * t["obs"].numel() / (a0.elapsed_time(a1) / 20 * 1e-3) / 1e12))
