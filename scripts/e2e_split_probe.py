"""Where the pipelined host-buffer step spends the host thread's time: wait (event + expansion on
the pool), policy, launch (the CUDA API calls of one range)."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
ne = 1 << 20
b = ph.HanabiBatch(dict(players=2), ne, (np.arange(ne) + 1).astype(np.int32))
D, A = b.obs_size, b.max_moves
pin = lambda *shape, dt: torch.empty(shape, dtype=dt).pin_memory()
obs, mask = pin(ne, D, dt=torch.uint8), pin(ne, A, dt=torch.uint8)
cur, rew, done, act = pin(ne, dt=torch.int32), pin(ne, dt=torch.int32), pin(ne, dt=torch.uint8), pin(ne, dt=torch.int32)
b.reset_host(obs, mask, cur)
launch, wait = b.bind_step_encode_host_range(act, rew, done, obs, mask, cur)
out = {}
for n_ranges in (8, 16, 32):
    rs = (ne // n_ranges + 127) // 128 * 128
    ranges = [(lo, min(lo + rs, ne) - lo) for lo in range(0, ne, rs)]
    pols = [ph.bind_host_random_legal_actions(mask[lo:lo + c], 7, act[lo:lo + c], first_game=lo) for lo, c in ranges]
    acc = [0.0, 0.0, 0.0]
    def step(t, timed):
        for (lo, c), pol in zip(ranges, pols):
            t0 = time.perf_counter(); wait(lo)
            t1 = time.perf_counter(); pol(t)
            t2 = time.perf_counter(); launch(lo, c)
            t3 = time.perf_counter()
            if timed:
                acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2
    for t in range(3): step(t, False)
    wait(-1)
    K = 30
    t0 = time.perf_counter()
    for t in range(K): step(3 + t, True)
    wait(-1)
    dt = (time.perf_counter() - t0) / K * 1e3
    out[n_ranges] = {"ms_per_step": dt, "wait_ms": acc[0] / K * 1e3, "policy_ms": acc[1] / K * 1e3,
                     "launch_ms": acc[2] / K * 1e3, "env_steps_per_s": ne / dt * 1e3}
print(json.dumps(out))
