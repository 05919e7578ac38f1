"""oracle/replay_oracle.py -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's replay memory for stack_size 1:
  OutOfGraphReplayMemory            agents/rainbow_copy/replay_memory.py:67-347
  OutOfGraphPrioritizedReplayMemory agents/rainbow_copy/prioritized_replay_memory.py:41-170
  SumTree                           agents/rainbow_copy/third_party/dopamine/sum_tree.py:30-205
PARITY PINNED: tests/test_replay_oracle.py runs it against the reference classes
themselves (imported with no-op stubs for the absent tensorflow/gin modules) where
/root/reference exists, and against tests/golden/replay_golden.npz recorded from them.
"""
import math

import numpy as np

DEFAULT_PRIORITY = 100.0


class SumTree(object):

  def __init__(self, capacity):
    depth = int(math.ceil(np.log2(capacity)))          # sum_tree.py:75
    self.nodes = [np.zeros(2 ** l) for l in range(depth + 1)]
    self.max_recorded_priority = 1.0

  def total(self):
    return self.nodes[0][0]

  def sample(self, query_value):                        # :99-141
    query_value = query_value * self.total()
    node = 0
    for level in self.nodes[1:]:
      left = node * 2
      if query_value < level[left]:
        node = left
      else:
        node = left + 1
        query_value -= level[left]
    return node

  def get(self, index):
    return self.nodes[-1][index]

  def set(self, index, value):                          # :181-205
    self.max_recorded_priority = max(value, self.max_recorded_priority)
    delta = value - self.nodes[-1][index]
    for level in reversed(self.nodes):
      level[index] += delta
      index //= 2


class ReplayOracle(object):

  def __init__(self, num_actions, observation_size, capacity, update_horizon=1, gamma=1.0,
               prioritized=True):
    self.A, self.D, self.cap, self.n = num_actions, observation_size, capacity, update_horizon
    self.disc = np.array([math.pow(gamma, k) for k in range(update_horizon)], dtype=np.float32)
    self.observations = np.zeros((capacity, observation_size), np.uint8)
    self.actions = np.zeros(capacity, np.int32)
    self.rewards = np.zeros(capacity, np.float32)
    self.terminals = np.zeros(capacity, np.uint8)
    self.legal_actions = np.zeros((capacity, num_actions), np.float32)
    self.add_count = 0
    self.tree = SumTree(capacity) if prioritized else None

  def cursor(self):
    return self.add_count % self.cap

  def is_full(self):
    return self.add_count >= self.cap

  def add(self, observation, action, reward, terminal, legal_actions, priority=DEFAULT_PRIORITY):
    c = self.cursor()                                   # replay_memory.py:155-164
    self.observations[c] = observation
    self.actions[c] = action
    self.rewards[c] = reward
    self.terminals[c] = terminal
    self.legal_actions[c] = legal_actions
    self.add_count += 1
    if self.tree is not None:
      self.tree.set(c, priority)                        # prioritized_replay_memory.py:101-107

  def is_valid_transition(self, index):                 # replay_memory.py:233-258, stack_size 1
    if index < 0 or index >= self.cap:
      return False
    if not self.is_full():
      if index >= self.cursor() - self.n:
        return False
    if index == (self.cursor() - 1) % self.cap:        # invalid_range
      return False
    return True

  def sample_transition_batch(self, indices):           # :290-347
    B = len(indices)
    state = self.observations[indices]
    action = self.actions[indices]
    reward = np.zeros(B, np.float32)
    terminal = np.zeros(B, np.uint8)
    next_state = np.zeros((B, self.D), np.uint8)
    next_legal = np.zeros((B, self.A), np.float32)
    for b, i in enumerate(indices):
      traj = [(i + j) % self.cap for j in range(self.n)]
      term = np.nonzero(self.terminals[traj])[0]
      if term.size == 0:
        reward[b] = self.disc.dot(self.rewards[traj])
      else:
        terminal[b] = 1
        t = int(term.min())
        reward[b] = self.disc[:t + 1].dot(self.rewards[traj[:t + 1]])
      boot = (i + self.n) % self.cap
      next_state[b] = self.observations[boot]
      next_legal[b] = self.legal_actions[boot]
    return state, action, reward, next_state, terminal, np.asarray(indices, np.int32), next_legal


def random_stream(rng, count, D, A, terminal_rate=0.1):
  """Synthetic transitions in the reference's formats."""
  obs = (rng.random((count, D)) < 0.4).astype(np.uint8)
  act = rng.integers(0, A, count).astype(np.int32)
  rew = rng.integers(-3, 2, count).astype(np.float32)
  term = (rng.random(count) < terminal_rate).astype(np.uint8)
  legal = np.where(rng.random((count, A)) < 0.5, 0.0, -np.inf).astype(np.float32)
  return obs, act, rew, term, legal
