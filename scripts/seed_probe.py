"""Set-up cost: NewBatch (seeding 2^19 mt19937 blocks) and the first BatchReset."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
torch.zeros(1, device="cuda:0"); torch.cuda.synchronize()
for n in (1 << 19, 1 << 20):
    seeds = (np.arange(n) + 1).astype(np.int32)
    for rep in range(2):
        t0 = time.perf_counter()
        b = ph.HanabiBatch(dict(players=2), n, seeds)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        b.reset(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print("games %d: NewBatch %.2f ms (incl. host seed list, malloc, H2D), first reset %.2f ms" % (n, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
        del b
