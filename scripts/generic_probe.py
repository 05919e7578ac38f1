"""Static (compile-time rules) vs generic (runtime rules) step kernel on the same workloads."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
n = 1 << 19
dev = torch.device("cuda:0")
def run(params, generic):
    b = ph.HanabiBatch(params, n, (np.arange(n) + 1).astype(np.int32))
    if generic: b.set_debug(0, True)
    obs = torch.empty((n, b.obs_size), dtype=torch.uint8, device=dev); mask = torch.empty((n, b.max_moves), dtype=torch.uint8, device=dev)
    cur = torch.empty(n, dtype=torch.int32, device=dev); rew = torch.empty(n, dtype=torch.int32, device=dev)
    done = torch.empty(n, dtype=torch.uint8, device=dev); act = torch.empty(n, dtype=torch.int32, device=dev)
    b.reset(); b.observe(obs, mask, cur)
    ph.select_action(None, mask, 1.0, seed=7, step=0, out=act)
    f = b.bind_step_encode_random(act, 7, rew, done, obs, mask, cur)
    for t in range(50): f(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for t in range(300): f(50 + t)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 300
cases = {"full2": dict(players=2), "full4": dict(players=4), "small2": dict(players=2, colors=2, ranks=5, hand_size=2, max_information_tokens=3, max_life_tokens=1),
         "small4": dict(players=4, colors=2, ranks=5, hand_size=2, max_information_tokens=3, max_life_tokens=1),
         "ppo_c3_h3_it3": dict(players=4, colors=3, ranks=5, hand_size=3, max_information_tokens=3, max_life_tokens=1)}
for name, p in cases.items():
    s, g = run(p, False), run(p, True)
    print("%-14s static %.4f ms  generic %.4f ms  (%.2fx)  %.3e steps/s static" % (name, s, g, g / s, n / s * 1e3))
