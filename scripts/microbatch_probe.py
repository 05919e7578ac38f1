"""Rainbow-policy rollout with the games cut into M micro-batches that share ONE set of
network buffers (input rows, hidden layer, logits): the buffers of a micro-batch fit the L2, are
overwritten by the next micro-batch before they are evicted, and never travel to HBM.
One CUDA graph per step; prints env-steps/s for several M."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
dev = torch.device("cuda:0")
n = 1 << 18
torch.manual_seed(0)
atoms, D0, A0 = 51, 658, 20
Kp, Np = (D0 + 31) // 32 * 32, (A0 * atoms + 63) // 64 * 64
tdt = torch.bfloat16
w1 = torch.zeros(Kp, 512, device=dev); w1[:D0] = torch.randn(D0, 512, device=dev) / D0 ** 0.5
w2 = torch.zeros(512, Np, device=dev); w2[:, :A0 * atoms] = torch.randn(512, A0 * atoms, device=dev) / 512 ** 0.5
w1, w2 = w1.to(tdt), w2.to(tdt)
b1, b2 = torch.zeros(512, device=dev, dtype=tdt), torch.zeros(Np, device=dev, dtype=tdt)
support = torch.linspace(-25.0, 25.0, atoms, device=dev)
out = {}
for M in (1, 4, 8, 16, 32):
    m = n // M
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    x = torch.empty((m, Kp), dtype=tdt, device=dev)
    logits = torch.empty((m, Np), dtype=tdt, device=dev)
    shards = []
    for i in range(M):
        b = ph.HanabiBatch(dict(players=2), m, (np.arange(m) + 1 + i * m).astype(np.int32))
        t = dict(mask=torch.empty((m, A0), dtype=torch.uint8, device=dev), cur=torch.empty(m, dtype=torch.int32, device=dev),
                 rew=torch.empty(m, dtype=torch.int32, device=dev), done=torch.empty(m, dtype=torch.uint8, device=dev),
                 act=torch.empty(m, dtype=torch.int32, device=dev))
        b.reset()
        sel = ph.bind_select_action_c51(logits, atoms, t["mask"], 0.0, seed=7, out=t["act"], support=support)
        shards.append((b, t, sel))
    # a micro-batch's turn: env step with the action chosen in its previous turn (the step kernel
    # writes the new observation's network rows into the shared x) -> MLP -> greedy action
    def turn(b, t, sel):
        b.step_encode_net(t["act"], x, t["rew"], t["done"], mask=t["mask"], cur_player=t["cur"])
        h2 = torch._addmm_activation(b1, x, w1)
        torch.addmm(b2, h2, w2, out=logits)
        sel(0)
    for b, t, sel in shards:  # first actions
        b.observe_net(x, mask=t["mask"], cur_player=t["cur"])
        h2 = torch._addmm_activation(b1, x, w1)
        torch.addmm(b2, h2, w2, out=logits)
        sel(0)
    def one_step():
        for b, t, sel in shards:
            turn(b, t, sel)
    for _ in range(3): one_step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
        one_step()
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 60
    e0.record()
    for _ in range(K): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out[M] = {"ms_per_step": ms, "env_steps_per_s": n / ms * 1e3, "games_per_micro_batch": m}
    del g, shards, x, logits
print(json.dumps(out))
