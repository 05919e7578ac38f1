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


# ---- HostExpandPacked: the host half of the packed link (hb_host.cc)
@pytest.mark.parametrize("dim", [106, 171, 281, 308, 640, 658, 956, 1041, 1280])
@pytest.mark.parametrize("n", [1, 2, 33, 4099])
def test_host_expand_packed_matches_unpackbits(ph, dim, n):
  import torch
  words = (dim + 127) // 128 * 4
  rng = np.random.default_rng(dim * 7 + n)
  bits = rng.integers(0, 2, size=(n, dim), dtype=np.uint8)
  padded = np.zeros((n, words * 32), dtype=np.uint8)
  padded[:, :dim] = bits
  padded[:, dim:] = rng.integers(0, 2, size=(n, words * 32 - dim), dtype=np.uint8)  # padding bits must not leak
  packed = np.packbits(padded, axis=1, bitorder="little").view(np.int32).reshape(n, words)
  t = torch.from_numpy(np.ascontiguousarray(packed))
  # dense rows at every alignment of the first byte (streaming stores need aligned spans)
  for shift in (0, 1, 2, 13, 31):
    flat = torch.full((n * dim + 64,), 7, dtype=torch.uint8)
    out = flat[shift:shift + n * dim].view(n, dim)
    ph.host_expand_packed(t, dim, out=out)
    assert (out.numpy() == bits).all()
    assert (flat[:shift] == 7).all() and (flat[shift + n * dim:] == 7).all()
  # padded rows
  wide = torch.full((n, dim + 22), 9, dtype=torch.uint8)
  ph.host_expand_packed(t, dim, out=wide)
  assert (wide[:, :dim].numpy() == bits).all() and (wide[:, dim:] == 9).all()


def test_host_expand_packed_large_dense_all_threads(ph):
  import torch
  dim, n = 658, 70000  # several 64 KB spans per worker
  rng = np.random.default_rng(5)
  padded = rng.integers(0, 2, size=(n, 768), dtype=np.uint8)
  packed = np.packbits(padded, axis=1, bitorder="little").view(np.int32).reshape(n, 24)
  out = ph.host_expand_packed(torch.from_numpy(np.ascontiguousarray(packed)), dim)
  assert (out.numpy() == padded[:, :dim]).all()
