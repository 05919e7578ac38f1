"""Small end-to-end case for compute-sanitizer: every kernel, ragged tile, several rule sets."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from hanabi_b200 import pyhanabi as ph
dev = torch.device("cuda:0")
for params in (dict(players=2), dict(players=5), dict(players=3, colors=4, ranks=3, hand_size=3, max_information_tokens=5, max_life_tokens=2),
               dict(players=4, random_start_player="true"), dict(players=3, hand_size=7, observation_type=2)):
    n = 203
    b = ph.HanabiBatch(params, n, list(range(1, n + 1)))
    b.enable_history()
    obs = torch.zeros((n, b.obs_size), dtype=torch.uint8, device=dev); mask = torch.zeros((n, b.max_moves), dtype=torch.uint8, device=dev)
    cur = torch.zeros(n, dtype=torch.int32, device=dev); rew = torch.zeros(n, dtype=torch.int32, device=dev)
    done = torch.zeros(n, dtype=torch.uint8, device=dev); act = torch.zeros(n, dtype=torch.int32, device=dev); sc = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.zeros((n, 1), dtype=torch.int32, device=dev)
    b.reset(); b.observe(obs, mask, cur)
    for t in range(40):
        ph.select_action(None, mask, 1.0, seed=3, step=t, out=act)
        if t % 2 and params.get("hand_size", 4) <= 5: b.agent_act(0 if b.num_players >= 2 and b.obs_size > 0 and params.get("hand_size", 4) <= 5 and not (params.get("hand_size") == 3 and False) else 1, 0b10, act, agent_state=st)
        b.step_encode(act, rew, done, obs, mask, cur, scores=sc)
    b.step(act, rew, done, sc, auto_reset=False); b.encode(obs, 0); b.legal_mask(mask); b.unpack_state()
    blob = b.get_state(); b.set_state(blob)
    h = [torch.zeros(x, dtype=d).pin_memory() for x, d in (((n, b.obs_size), torch.uint8), ((n, b.max_moves), torch.uint8), ((n,), torch.int32), ((n,), torch.int32), ((n,), torch.uint8), ((n,), torch.int32))]
    b.reset_host(h[0], h[1], h[2])
    ph.host_random_legal_actions(h[1], 1, 0, out=h[5])
    b.step_encode_host(h[5], h[3], h[4], h[0], h[1], h[2])
    launch, wait = b.bind_step_encode_host_range(h[5], h[3], h[4], h[0], h[1], h[2])
    launch(0, 128); launch(128, n - 128); wait(-1)
    # random-policy fusion, packed rows, network-input rows, error flags, single-game export
    b.reset(); b.observe(obs, mask, cur)
    ph.select_action(None, mask, 1.0, seed=3, step=0, out=act)
    for t in range(1, 12):
        b.step_encode_random(act, 3, t, rew, done, obs, mask, cur)
    bits = torch.zeros((n, b.packed_words), dtype=torch.int32, device=dev)
    b.step_encode_packed(act, rew, done, bits, mask, cur)
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        x = torch.zeros((n, (b.obs_size + 31) // 32 * 32 + 8), dtype=dt, device=dev)
        ph.select_action(None, mask, 1.0, seed=3, step=50, out=act)
        b.step_encode_net(act, x, rew, done, obs_bits=bits, mask=mask, cur_player=cur)
        b.observe_net(x)
        ph.expand_packed(bits, b.obs_size, x)
    b.error_flags(); b.export_state(n - 1)
    # recorder + replay memory
    from hanabi_b200 import actor, replay
    mem = replay.DeviceReplayMemory(b.max_moves, b.obs_size, 4096, update_horizon=3, gamma=0.99)
    rec = actor.DeviceRecorder(n, b.num_players, b.obs_size, b.max_moves, mem)
    b.reset(); b.observe_packed(bits, mask, cur)
    for t in range(60):
        ph.select_action(None, mask, 1.0, seed=5, step=t, out=act)
        rec.before_step(cur, bits, mask, act)
        b.step_encode_packed(act, rew, done, bits, mask, cur)
        rec.after_step(rew, done, return_count=(t % 7 == 0))
    if mem.add_count > 40:
        idx = mem.sample_index_batch(32)
        mem.sample_transition_batch(indices=idx); mem.set_priority(idx, torch.rand(32, device=dev) + 0.1)
z = torch.linspace(-25, 25, 51, device=dev)
lg = torch.randn(37, 20, 51, device=dev); m = torch.ones(37, 20, dtype=torch.uint8, device=dev)
ph.select_action(lg, m, 0.1, 1, 2, support=z); ph.c51_q_values(lg, z, m)
for dt in (torch.float32, torch.bfloat16, torch.float16):
    pad = torch.zeros((37, 1024), dtype=dt, device=dev); pad[:, :1020] = lg.view(37, 1020).to(dt)
    ph.select_action_c51(pad, 51, m, 0.1, 1, 2, support=z)
    ph.sample_masked_categorical(pad[:, :32].contiguous(), m, 1, 2)
ph.project_distribution(torch.randn(33, 51, device=dev), torch.rand(33, 51, device=dev), z)
ph.c51_target_distribution(torch.randn(37, device=dev), torch.zeros(37, dtype=torch.uint8, device=dev), lg, m, z, 0.95)
torch.cuda.synchronize(); print("sanitize case ok")
