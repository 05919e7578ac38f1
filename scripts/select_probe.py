"""Device time of the random-policy select kernel alone and of an empty step (launch gaps)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
n = 1 << 20
dev = torch.device("cuda:0")
b = ph.HanabiBatch(dict(players=2), n, (np.arange(n) + 1).astype(np.int32))
obs = torch.empty((n, b.obs_size), dtype=torch.uint8, device=dev); mask = torch.empty((n, b.max_moves), dtype=torch.uint8, device=dev)
cur = torch.empty(n, dtype=torch.int32, device=dev); rew = torch.empty(n, dtype=torch.int32, device=dev)
done = torch.empty(n, dtype=torch.uint8, device=dev); act = torch.empty(n, dtype=torch.int32, device=dev)
b.reset(); b.observe(obs, mask, cur)
sel = ph.bind_select_action(None, mask, 1.0, seed=7, out=act)
step = b.bind_step_encode(act, rew, done, obs, mask, cur)
for i in range(50): sel(i); step()
torch.cuda.synchronize()
def timed(fn, k=200):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(k): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3
print("select alone, back to back: %.1f us" % timed(lambda i: sel(i)))
print("step alone, back to back:   %.1f us" % timed(lambda i: step()))
print("select + step:              %.1f us" % timed(lambda i: (sel(i), step())))
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def cold(i):
    junk.zero_(); sel(i)
print("zero 256MB + select:        %.1f us" % timed(cold, 50))
print("zero 256MB alone:           %.1f us" % timed(lambda i: junk.zero_(), 50))
