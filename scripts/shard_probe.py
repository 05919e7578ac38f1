"""Does splitting one GPU's games into S independent shards on S streams hide the kernel tail?"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
N = 1 << 20
dev = torch.device("cuda:0")
def run(S, K=1000, W=100):
    n = N // S
    shards = []
    for s in range(S):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            b = ph.HanabiBatch(dict(players=2), n, (np.arange(n) + 1 + s * n).astype(np.int32))
            obs = torch.empty((n, b.obs_size), dtype=torch.uint8, device=dev); mask = torch.empty((n, b.max_moves), dtype=torch.uint8, device=dev)
            cur = torch.empty(n, dtype=torch.int32, device=dev); rew = torch.empty(n, dtype=torch.int32, device=dev)
            done = torch.empty(n, dtype=torch.uint8, device=dev); act = torch.empty(n, dtype=torch.int32, device=dev)
            b.reset(); b.observe(obs, mask, cur)
            ph.select_action(None, mask, 1.0, seed=7, step=0, out=act)
            f = b.bind_step_encode_random(act, 7, rew, done, obs, mask, cur, stream=st)
        shards.append((b, st, f, (obs, mask, cur, rew, done, act)))
    torch.cuda.synchronize()
    for t in range(W):
        for b, st, f, _ in shards: f(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for b, st, f, _ in shards: st.wait_event(e0)
    for t in range(K):
        for b, st, f, _ in shards: f(W + t)
    for b, st, f, _ in shards: main.wait_stream(st)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("shards %d x %d games: %.4f ms/step  %.3e steps/s" % (S, n, ms, N / ms * 1e3))
for S in (1, 2, 3, 4, 8):
    run(S)
