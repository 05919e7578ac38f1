"""e2e probe: the step kernel writing its outputs straight into pinned (mapped) HOST memory
against the staged form (kernel -> device staging -> cudaMemcpyAsync).  Ranges are emulated
with independent batches, one stream each; the host random-legal policy runs inside the loop."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
from hanabi_b200.pyhanabi import ffi, lib

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 8
params = dict(players=2)
pin = lambda *shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
n = ne // nr


class Range:
  def __init__(self, base):
    self.b = ph.HanabiBatch(params, n, (np.arange(n) + base).astype(np.int32))
    D, A = self.b.obs_size, self.b.max_moves
    self.obs, self.mask = pin(n, D, dt=torch.uint8), pin(n, A, dt=torch.uint8)
    self.cur, self.rew = pin(n, dt=torch.int32), pin(n, dt=torch.int32)
    self.done, self.act = pin(n, dt=torch.uint8), pin(n, dt=torch.int32)
    self.d_act = torch.zeros(n, dtype=torch.int32, device="cuda")
    self.b.reset_host(self.obs, self.mask, self.cur)
    self.stream = torch.cuda.Stream()
    c = lambda t, ty: ffi.cast(ty, t.data_ptr())
    self.args_zero = (c(self.act, "int32_t*"), c(self.rew, "int32_t*"), c(self.done, "uint8_t*"), ffi.NULL,
                      c(self.obs, "uint8_t*"), D, c(self.mask, "uint8_t*"), c(self.cur, "int32_t*"), 1)
    self.args_dact = (c(self.d_act, "int32_t*"),) + self.args_zero[1:]

  def step_zero(self, t, dev_actions):
    self.stream.synchronize()
    ph.host_random_legal_actions(self.mask, seed=7, step=t, out=self.act)
    with torch.cuda.stream(self.stream):
      if dev_actions:
        self.d_act.copy_(self.act, non_blocking=True)
        args = self.args_dact
      else:
        args = self.args_zero
      rc = lib.BatchStepEncodeScores(self.b._c, *args, ffi.cast("void*", self.stream.cuda_stream))
      assert rc == 0


def run(label, fn, k=36):
  for t in range(3):
    fn(t)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for t in range(k):
    fn(3 + t)
  torch.cuda.synchronize()
  ms = (time.perf_counter() - t0) / k * 1e3
  print("%-60s %.3f ms/step  %.3e steps/s  %.1f GB/s down" % (label, ms, ne / ms * 1e3, ne * 687 / ms / 1e6))


rs = [Range(1 + i * n) for i in range(nr)]
for dev_actions in (True, False):
  def zero(t):
    for r in rs:
      r.step_zero(t, dev_actions)
  run("kernel stores into host memory, %d ranges, actions %s" % (nr, "uploaded" if dev_actions else "read from host"), zero)

# staged reference: the shipped ranged call
hb = ph.HanabiBatch(params, ne, (np.arange(ne) + 1).astype(np.int32))
D, A = hb.obs_size, hb.max_moves
h_obs, h_mask = pin(ne, D, dt=torch.uint8), pin(ne, A, dt=torch.uint8)
h_cur, h_rew = pin(ne, dt=torch.int32), pin(ne, dt=torch.int32)
h_done, h_act = pin(ne, dt=torch.uint8), pin(ne, dt=torch.int32)
hb.reset_host(h_obs, h_mask, h_cur)
launch, wait = hb.bind_step_encode_host_range(h_act, h_rew, h_done, h_obs, h_mask, h_cur)
rsz = (ne // nr + 127) // 128 * 128
ranges = [(lo, min(lo + rsz, ne) - lo) for lo in range(0, ne, rsz)]
views = [(h_mask[lo:lo + cnt], h_act[lo:lo + cnt]) for lo, cnt in ranges]


def staged(t):
  for (lo, cnt), (m_v, a_v) in zip(ranges, views):
    wait(lo)
    ph.host_random_legal_actions(m_v, seed=7, step=t, out=a_v)
    launch(lo, cnt)


run("staged (BatchStepEncodeHostRange), %d ranges" % len(ranges), staged)
wait(-1)
# transfers alone: how fast is the link with one big copy
d = torch.empty(ne * 687, dtype=torch.uint8, device="cuda")
h = torch.empty(ne * 687, dtype=torch.uint8).pin_memory()
for _ in range(3):
  h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
  h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 10 * 1e3
print("one cudaMemcpyAsync D2H of %d MB: %.3f ms  %.1f GB/s" % (ne * 687 >> 20, ms, ne * 687 / ms / 1e6))
