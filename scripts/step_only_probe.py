"""Runs BatchStep (no observation output) + BatchObserve separately, for profiling the step part."""
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
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 320):
    sel(i); b.step(act, rew, done); b.legal_mask(mask)
torch.cuda.synchronize(); print("ok")
