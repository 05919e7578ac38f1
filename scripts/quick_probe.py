"""Quick device timing probe (not the bench): random-legal rollout, N games."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph

def run(params, n, steps=100, warm=20):
    dev = torch.device("cuda:0")
    seeds = (np.arange(n) + 1).astype(np.int32)
    t0 = time.time()
    b = ph.HanabiBatch(params, n, seeds)
    obs = torch.empty((n, b.obs_size), dtype=torch.uint8, device=dev)
    mask = torch.empty((n, b.max_moves), dtype=torch.uint8, device=dev)
    cur = torch.empty(n, dtype=torch.int32, device=dev)
    rew = torch.empty(n, dtype=torch.int32, device=dev)
    done = torch.empty(n, dtype=torch.uint8, device=dev)
    act = torch.empty(n, dtype=torch.int32, device=dev)
    b.reset(); b.observe(obs, mask, cur)
    torch.cuda.synchronize()
    t_init = time.time() - t0
    def loop(k, base):
        for t in range(k):
            ph.select_action(None, mask, 1.0, seed=7, step=base + t, out=act)
            b.step_encode(act, rew, done, obs, mask, cur)
    loop(warm, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); loop(steps, warm); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    sps = n * steps / (ms * 1e-3)
    # step kernel alone
    e0.record()
    for t in range(steps): b.step_encode(act, rew, done, obs, mask, cur)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1)
    st = b.read_stats()
    print(f"{params} n={n} init={t_init:.2f}s  rollout {sps/1e6:.1f} M steps/s ({ms/steps*1e3:.0f} us/step); "
          f"step kernel alone {n*steps/(ms2*1e-3)/1e6:.1f} M/s; episodes={st['episodes']} mean_len={st['length_sum']/max(st['episodes'],1):.1f}")

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for n in (1 << 16, 1 << 20):
        run(dict(players=2), n)
    run(dict(players=5), 1 << 20)
    run(dict(players=2, colors=2, ranks=5, hand_size=2, max_information_tokens=3, max_life_tokens=1), 1 << 20)
