"""Host side of the random-legal policy (hb_host.cc): the persistent worker pool and the
counter-based draw keyed (seed, step, global game index) -- checked against an independent
numpy restatement of hb_policy_rng.h (bench.policy_draw, the checker of the bench's
cross-rank parity leg)."""
import threading

import numpy as np
import pytest

import bench


@pytest.fixture(scope="module")
def ph(built_lib):
  from hanabi_b200 import pyhanabi
  assert pyhanabi.try_cdef() and pyhanabi.try_load()
  return pyhanabi


def _mask(n, A, seed):
  rng = np.random.default_rng(seed)
  m = (rng.random((n, A)) < 0.35).astype(np.uint8)
  m[n // 3] = 0          # a row without a legal move -> -1
  m[n // 2] = 1
  return m


@pytest.mark.parametrize("A", [10, 11, 20, 38, 48, 64])
def test_host_policy_matches_numpy_twin(ph, A):
  import torch
  n = 20000
  m = _mask(n, A, A)
  for first in (0, 12345, (1 << 33) + 7):
    out = ph.host_random_legal_actions(torch.from_numpy(m), seed=7, step=99, first_game=first)
    ref = bench.policy_draw(m, 7, 99, np.arange(n, dtype=np.int64) + first)
    assert (out.numpy() == ref).all()
  assert out[n // 3] == -1


def test_host_policy_ranges_equal_whole(ph):
  """Ranges of a batch with their first_game offsets draw what one call over the whole
  batch draws (what the pipelined host step relies on)."""
  import torch
  n, A = 50000, 20
  m = torch.from_numpy(_mask(n, A, 3))
  whole = ph.host_random_legal_actions(m, seed=5, step=17, first_game=1000)
  parts = torch.empty(n, dtype=torch.int32)
  for lo in range(0, n, 6400):
    hi = min(n, lo + 6400)
    run = ph.bind_host_random_legal_actions(m[lo:hi], 5, parts[lo:hi], first_game=1000 + lo)
    run(17)
  assert torch.equal(whole, parts)
  # non-zero bytes other than 1 count as legal
  m2 = m.clone(); m2[m2 != 0] = 200
  assert torch.equal(ph.host_random_legal_actions(m2, seed=5, step=17, first_game=1000), whole)


def test_host_pool_survives_concurrent_callers(ph):
  import torch
  n, A = 30000, 20
  m = torch.from_numpy(_mask(n, A, 9))
  want = [ph.host_random_legal_actions(m, seed=1, step=s).clone() for s in range(4)]
  errors = []

  def worker(s):
    for _ in range(25):
      got = ph.host_random_legal_actions(m, seed=1, step=s)
      if not torch.equal(got, want[s]):
        errors.append(s)

  threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
  for t in threads: t.start()
  for t in threads: t.join()
  assert not errors
  assert ph.host_policy_threads() >= 1
