"""Where does the host-buffer step's time go?  Transfers alone (no-op actions), with the host policy,
and with two half-batches whose policy and transfers alternate (double buffering)."""
import sys, time, threading
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
params = dict(players=2)
pin = lambda *shape, dt: torch.empty(shape, dtype=dt).pin_memory()
class Host:
    def __init__(self, n, base):
        self.n = n
        self.b = ph.HanabiBatch(params, n, (np.arange(n) + base).astype(np.int32))
        D, A = self.b.obs_size, self.b.max_moves
        self.obs, self.mask = pin(n, D, dt=torch.uint8), pin(n, A, dt=torch.uint8)
        self.cur, self.rew = pin(n, dt=torch.int32), pin(n, dt=torch.int32)
        self.done, self.act = pin(n, dt=torch.uint8), pin(n, dt=torch.int32)
        self.b.reset_host(self.obs, self.mask, self.cur)
    def policy(self, t): ph.host_random_legal_actions(self.mask, seed=7, step=t, out=self.act)
    def step(self): self.b.step_encode_host(self.act, self.rew, self.done, self.obs, self.mask, self.cur)
def timeit(fn, k=12):
    for t in range(3): fn(t)
    t0 = time.perf_counter()
    for t in range(k): fn(3 + t)
    return (time.perf_counter() - t0) / k * 1e3
h = Host(ne, 1)
noop = torch.full((ne,), -1, dtype=torch.int32).pin_memory()
def only_transfer(t): h.b.step_encode_host(noop, h.rew, h.done, h.obs, h.mask, h.cur)
ms = timeit(only_transfer); print("transfers only (no-op actions): %.2f ms  %.1f GB/s" % (ms, ne * 687 / ms / 1e6))
def only_policy(t): h.policy(t)
print("host policy only: %.2f ms" % timeit(only_policy))
def both(t): h.policy(t); h.step()
ms = timeit(both); print("policy + step, serial: %.2f ms  %.3e steps/s" % (ms, ne / ms * 1e3))
# double buffering: two halves; policy of one half runs on a thread while the other half transfers
a, b2 = Host(ne // 2, 1), Host(ne // 2, 1 + ne // 2)
def pipelined(k):
    a.policy(0)
    t0 = time.perf_counter()
    for t in range(k):
        th = threading.Thread(target=b2.policy, args=(t,)); th.start()
        a.step()
        th.join()
        th = threading.Thread(target=a.policy, args=(t + 1,)); th.start()
        b2.step()
        th.join()
    return (time.perf_counter() - t0) / k * 1e3
pipelined(3)
ms = pipelined(12); print("two halves, policy || transfer: %.2f ms  %.3e steps/s" % (ms, ne / ms * 1e3))
