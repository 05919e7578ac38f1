"""oracle/actor_oracle.py -- TEST INFRASTRUCTURE ONLY.

Restatement of the per-episode transition bookkeeping of the reference's DQN/Rainbow
agent and episode loop for ONE game (agents/rainbow_copy/dqn_agent.py:286-385,
agents/run_experiment.py:344-403): begin_episode/step record (reward, o, l, a);
end_episode posts, per player, (o_t, a_t, r_{t+1}, terminal, l_t) consecutively.
Pinned (tests/test_actor_golden.py) by tests/golden/actor_golden.npz: the stream the
reference's own DQNAgent methods posted when the reference's run_one_episode drove them
on seeded games (tests/golden/make_actor_golden.py; the agent object is created without
its TF-graph-building constructor, its three network hooks are scripted).
"""
import numpy as np


class EpisodeBook(object):

  def __init__(self, num_players):
    self.P = num_players
    self.transitions = [[] for _ in range(num_players)]      # dqn_agent.py:181
    self.since = np.zeros(num_players, np.float32)           # reward_since_last_action

  def act(self, player, observation, legal_mask, action):
    # run_experiment.py:383-397: step(reward_since_last_action[p], ...) or begin_episode(...)
    reward = float(self.since[player]) if self.transitions[player] else 0.0
    self.transitions[player].append((reward, np.array(observation, np.uint8), np.array(legal_mask, np.uint8),
                                     int(action)))
    self.since[player] = 0                                   # :399-400

  def env_result(self, reward, done):
    self.since += reward                                     # :374-377
    if not done:
      return None
    posted = []                                              # _post_transitions :359-385
    for player in range(self.P):
      tr = self.transitions[player]
      for index, (_, obs, legal, action) in enumerate(tr):
        final = index == len(tr) - 1
        r = float(self.since[player]) if final else tr[index + 1][0]
        posted.append((obs.tobytes(), action, np.float32(r), int(final), legal.tobytes()))
      self.transitions[player] = []
    self.since[:] = 0
    return posted
