"""Device replay memory micro-benchmark: add / sample / gather / set_priority (CUDA events)."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import replay
dev = torch.device("cuda:0")
D, A, cap, n = 658, 20, 1 << 22, 5
mem = replay.DeviceReplayMemory(A, D, cap, update_horizon=n, gamma=0.99, seed=1)
B = 1 << 20
obs = (torch.rand((B, D), device=dev) < 0.4).to(torch.uint8)
act = torch.randint(0, A, (B,), device=dev, dtype=torch.int32)
rew = torch.zeros(B, device=dev); term = (torch.rand(B, device=dev) < 0.05).to(torch.uint8)
legal = torch.zeros((B, A), device=dev)
def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_add = timed(lambda: mem.add(obs, act, rew, term, legal), 8)
S = 1 << 15
idx = mem.sample_index_batch(S)
t_sample = timed(lambda: mem.sample_index_batch(S), 20)
t_gather = timed(lambda: mem.sample_transition_batch(indices=idx), 20)
pri = torch.rand(S, device=dev)
t_setp = timed(lambda: mem.set_priority(idx, pri), 20)
bytes_add = B * (D + 4 * A + 9) * 2
bytes_gather = S * (2 * D + 4 * A + 9) * 2
peak = 6544.3
out = {"capacity": cap, "obs_dim": D, "num_actions": A, "update_horizon": n,
       "add": {"transitions": B, "ms": t_add, "transitions_per_s": B / t_add * 1e3, "GB_per_s": bytes_add / t_add / 1e6, "frac_of_measured_hbm": bytes_add / t_add / 1e6 / peak},
       "sample_index_batch": {"batch": S, "ms": t_sample},
       "sample_transition_batch": {"batch": S, "ms": t_gather, "GB_per_s": bytes_gather / t_gather / 1e6},
       "set_priority": {"batch": S, "ms": t_setp}}
print(json.dumps(out))
del mem
