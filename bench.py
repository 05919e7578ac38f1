#!/usr/bin/env python
"""bench.py -- Hanabi env-steps/sec (step + encode + legal mask) on B200.

A "step" is one pass of the hot path over one batch of games: every game takes
one move of a uniform-random-legal policy (BatchSelectAction, counter-based RNG
keyed (7, step, game)), the move is applied, cards are dealt from the game's own
mt19937 stream, finished games are re-dealt (auto-reset), and the next player's
canonical observation (uint8 [games, obs_dim]), legal mask, reward, done flag
and player index are written (BatchStepEncode).  Workload = BASELINE.json
configs[0]/north_star: Hanabi-Full 2-player, 658-dim observation.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--games G] [--impl reference]

One process per GPU; under torchrun the ranks shard the games (no collective on
the step path) and one tiny NCCL all-reduce aggregates the episode statistics.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "hanabi_full_2p": dict(players=2),
    "hanabi_full_3p": dict(players=3),
    "hanabi_full_4p": dict(players=4),
    "hanabi_full_5p": dict(players=5),
    "hanabi_small_2p": dict(players=2, colors=2, ranks=5, hand_size=2, max_information_tokens=3,
                            max_life_tokens=1),
    # the PPO / REINFORCE scripts' default (training/hanabi_small/train_ppo_agent_small_env.py:111-112)
    "hanabi_small_4p": dict(players=4, colors=2, ranks=5, hand_size=2, max_information_tokens=3,
                            max_life_tokens=1),
}
# Algorithmic bytes per env-step (SURVEY.md section 8d / BASELINE.md section 4):
# obs_dim + A (uint8 mask) + 4 (reward) + 1 (done) + 4 (action) + 2 * S (packed state r+w).
ALGO_BYTES = {"hanabi_full_2p": 815, "hanabi_full_3p": 1155, "hanabi_full_4p": 1248, "hanabi_full_5p": 1529,
              "hanabi_small_2p": 255, "hanabi_small_4p": 411}
POLICY_SEED = 7
INIT_POLICY_STEP = 1 << 40  # key of the rollout's first actions (the steps count from 0)
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=2000)
  ap.add_argument("--warmup", type=int, default=100)
  ap.add_argument("--games", type=int, default=1 << 20, help="games per GPU")
  ap.add_argument("--workload", default="hanabi_full_2p", choices=sorted(WORKLOADS))
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--e2e-steps", type=int, default=12)
  ap.add_argument("--e2e-games", type=int, default=0,
                  help="games of the host-buffer leg (0 = the same games per GPU as the device leg)")
  ap.add_argument("--cpu-seconds", type=float, default=8.0)
  ap.add_argument("--reference-seconds", type=float, default=3.0,
                  help="--impl reference: minimum wall time of the timed CPU rollout")
  ap.add_argument("--python-seconds", type=float, default=2.0,
                  help="seconds per Python-path CPU leg (BASELINE.md section 3, C1-C3); 0 = skip")
  ap.add_argument("--no-parity-check", action="store_true",
                  help="skip the replay of sampled games on the CPU reference after the timed region")
  ap.add_argument("--parity-games", type=int, default=64)
  ap.add_argument("--no-policy-extras", action="store_true",
                  help="skip the short Rainbow-policy rollout reported under `policies`")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--separate-policy", action="store_true",
                  help="random policy as its own kernel (BatchSelectAction) instead of the step "
                       "kernel's fused random-action output (BatchStepEncodeRandom)")
  ap.add_argument("--mlp-dtype", choices=["fp32", "tf32", "bf16", "fp16"], default="tf32",
                  help="element type of the Rainbow policy's two library GEMMs (accumulation is fp32)")
  ap.add_argument("--record", action="store_true",
                  help="rainbow policy: run the whole actor loop -- the device recorder keeps every "
                       "seat's episode and appends finished episodes to a device replay memory")
  ap.add_argument("--replay-capacity", type=int, default=1 << 22)
  ap.add_argument("--graph-steps", type=int, default=0,
                  help="rainbow / adhoc (all rows) policies: env steps captured per CUDA graph "
                       "(0 = 1, or 8 below 32768 games per GPU where the step is launch-bound)")
  ap.add_argument("--no-graph", action="store_true",
                  help="rainbow / adhoc (all rows) policies: launch every kernel of a step from the "
                       "host instead of replaying one captured CUDA graph per step")
  ap.add_argument("--adhoc-all-rows", action="store_true",
                  help="ad-hoc team: run the Rainbow MLP on every game instead of only those whose "
                       "player to act is the Rainbow seat")
  ap.add_argument("--shards", type=int, default=2,
                  help="random policy only: split each GPU's games into this many independent "
                       "batches on their own CUDA streams, so that one launch's last, partly "
                       "filled wave of CTAs overlaps the other's body (1 = one launch per step)")
  ap.add_argument("--packed", action="store_true", default=True, help=argparse.SUPPRESS)
  ap.add_argument("--no-packed", dest="packed", action="store_false",
                  help="skip the bit-packed observation output (the `packed_obs` extra key)")
  ap.add_argument("--python-baselines", action="store_true", help=argparse.SUPPRESS)  # now the default
  ap.add_argument("--cpu-python-worker", default=None, help=argparse.SUPPRESS)
  ap.add_argument("--policy", default="random", choices=["random", "rainbow", "adhoc", "ppo"],
                  help="random: uniform legal (headline); rainbow: fixed-weight C51 MLP "
                       "obs->512->A*51 (torch GEMMs) + fused masked argmax, epsilon 0; adhoc: "
                       "seat 0 Rainbow, other seats the batched RuleBasedAgent sharing one object "
                       "(training/run_adhoc_experiment.py team 1)")
  return ap.parse_args()


# ------------------------------------------------------------ clock sampling
class ClockSampler(object):
  """Samples SM clock and throttle reasons through NVML while the timed region runs."""

  def __init__(self, index):
    self.samples = []
    self.reasons = set()
    self.max_mhz = None
    self._stop = threading.Event()
    self._thread = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
    except Exception:  # pylint: disable=broad-except
      self.nv = None

  def _sample(self):
    nv = self.nv
    try:
      self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
      r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
      names = {
          "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
          "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
          "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
          "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
          "hw_power_brake_slowdown": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
      }
      for name, bit in names.items():
        if r & bit:
          self.reasons.add(name)
    except Exception:  # pylint: disable=broad-except
      pass

  def start(self):
    if self.nv is None:
      return
    def run():
      while not self._stop.is_set():
        self._sample()
        time.sleep(0.004)
    self._thread = threading.Thread(target=run, daemon=True)
    self._thread.start()

  def stop(self):
    if self.nv is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    self._sample()
    self._stop.set()
    if self._thread:
      self._thread.join()
    s = sorted(self.samples)
    return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------- CPU baseline
def cpu_reference_rate(workload, seconds, threads=None):
  """Times the reference's own CPU engine (oracle/_ref, unmodified sources) -- or
  the C port when it was not built -- on this box's host cores: random-legal
  rollout with auto-reset, current-player observation + legal mask per step.
  One long threaded rollout of at least `seconds` wall time (thread start-up and
  cache warm-up amortised), preceded by an untimed calibration pass."""
  import numpy as np
  from oracle import oracle as O
  threads = threads or os.cpu_count() or 1
  kind = "ref" if O.have_ref() else "oracle"
  if kind == "oracle":
    threads = 1
  n = 512 * threads
  b = O.CpuBatch(kind, WORKLOADS[workload], n, (np.arange(n) + 1).astype(np.int32))
  b.reset()
  b.rollout(4, threads)  # page in, desynchronise the games a little
  t0 = time.perf_counter()
  b.rollout(16, threads)  # calibration
  per_pass = max((time.perf_counter() - t0) / 16, 1e-6)
  steps = max(16, int(seconds / per_pass) + 1)
  t0 = time.perf_counter()
  total, episodes, score_sum = b.rollout(steps, threads)
  dt = time.perf_counter() - t0
  return {"value": total / dt, "unit": "env-steps/s", "cores": threads,
          "kind": "reference" if kind == "ref" else "port",
          "sample": "%d games x %d steps, random-legal policy, auto-reset, %.1f s" % (n, steps, dt),
          "seconds": dt, "games": n, "steps": steps}


def python_worker(mode, seconds, seed, workload):
  """Child process (HANABI_B200_LIB = the reference's own libpyhanabi.so from
  oracle/_ref): the Python layers a reference user runs today, on the reference's
  native engine.  mode "rl_env": HanabiEnv.reset/step with the full per-player
  dicts (what every trainer calls, rl_env.py:343-366); mode "lean": apply_move +
  deal + current-player observation + encode + legal uids."""
  import random
  from hanabi_b200 import pyhanabi, rl_env
  cfg = dict(WORKLOADS[workload], seed=seed)
  rnd = random.Random(seed)
  steps = 0
  t0 = time.perf_counter()
  if mode == "rl_env":
    env = rl_env.HanabiEnv(cfg)
    obs = env.reset()
    while time.perf_counter() - t0 < seconds:
      legal = obs["player_observations"][obs["current_player"]]["legal_moves_as_int"]
      obs, _, done, _ = env.step(rnd.choice(legal))
      steps += 1
      if done:
        obs = env.reset()
  else:
    game = pyhanabi.HanabiGame(cfg)
    enc = pyhanabi.ObservationEncoder(game)
    state = None
    while time.perf_counter() - t0 < seconds:
      if state is None or state.is_terminal():
        state = game.new_initial_state()
        while state.cur_player() == pyhanabi.CHANCE_PLAYER_ID:
          state.deal_random_card()
      moves = state.legal_moves()
      state.apply_move(rnd.choice(moves))
      while state.cur_player() == pyhanabi.CHANCE_PLAYER_ID:
        state.deal_random_card()
      if not state.is_terminal():
        enc.encode(state.observation(state.cur_player()))
        [game.get_move_uid(m) for m in state.legal_moves()]
      steps += 1
  print(json.dumps({"steps": steps, "seconds": time.perf_counter() - t0}), flush=True)


def python_baselines(workload, seconds=2.0):
  """BASELINE.md section 3, C1-C3, on this box's host cores."""
  import subprocess
  from oracle import oracle as O
  if not os.path.exists(O.REF_PYHANABI_SO):
    return None
  env = dict(os.environ, HANABI_B200_LIB=O.REF_PYHANABI_SO)
  cores = os.cpu_count() or 1

  def launch(mode, count):
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-python-worker",
                               "%s:%g:%d" % (mode, seconds, 1 + i), "--workload", workload],
                              env=env, stdout=subprocess.PIPE, text=True) for i in range(count)]
    outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
    return sum(o["steps"] for o in outs) / max(o["seconds"] for o in outs)

  return {"unit": "env-steps/s", "cores": cores, "seconds_each": seconds,
          "rl_env_loop_1_process": launch("rl_env", 1), "lean_pyhanabi_loop_1_process": launch("lean", 1),
          "rl_env_loop_1_process_per_core": launch("rl_env", cores),
          "lean_pyhanabi_loop_1_process_per_core": launch("lean", cores),
          "note": "reference libpyhanabi.so (oracle/_ref) driven by the Python binding/rl_env layer"}


def run_reference_arm(args, rank):
  """--impl reference: the reference's CPU implementation of the path, all host
  threads.  The driver's --steps/--warmup are echoed; the timed rollout itself runs
  for at least --reference-seconds of wall time (a 20-step sample is 40 ms: thread
  start-up and cold caches would dominate it), so `ms_per_step` is the steady-state
  time of one step over `games_per_step` games."""
  if rank != 0:
    return
  r = cpu_reference_rate(args.workload, max(args.reference_seconds, 2.0))
  value = r["value"]
  sample = "%d games per step, %d steps timed in one threaded rollout of %.1f s (bounded sample of %s), random-legal policy" % (
      r["games"], r["steps"], r["seconds"], args.workload)
  line = {
      "impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s",
      "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
      "ms_per_step": r["seconds"] / r["steps"] * 1e3, "higher_is_better": True, "scaling": "weak",
      "vs_baseline": None, "dtype": "u8", "data": "synthetic",
      "config": {"workload": args.workload, "games_per_step": r["games"], "steps_timed": r["steps"],
                 "policy": "uniform random legal",
                 "note": "reference C++ engine (oracle/_ref) on host cores; timed for >= %.0f s "
                         "whatever --steps says" % max(args.reference_seconds, 2.0)},
      "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": r["cores"],
                       "kind": r["kind"], "sample": sample},
      "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line), flush=True)


# ------------------------------------------------ parity of sampled games (checker)
def _mix64(z):
  import numpy as np
  with np.errstate(over="ignore"):
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def policy_draw(mask, seed, step, games):
  """numpy twin of the engine's counter-based uniform-legal draw
  (hanabi_b200/csrc/hb_policy_rng.h): action of every row of `mask` for the key
  (seed, step, games[i]); -1 where nothing is legal."""
  import numpy as np
  with np.errstate(over="ignore"):
    inner = _mix64(np.uint64(seed) ^ (np.uint64(step) * np.uint64(0xD1B54A32D192ED03)))
    bits = _mix64(inner ^ (games.astype(np.uint64) * np.uint64(0x8CB92BA72F3D8DD7)))
  n_legal = mask.sum(axis=1).astype(np.uint64)
  k = (((bits & np.uint64(0xffffffff)) * n_legal) >> np.uint64(32)).astype(np.int64)
  order = np.cumsum(mask != 0, axis=1) - 1  # rank of each legal move within its row
  hit = (mask != 0) & (order == k[:, None])
  act = np.where(hit.any(axis=1), hit.argmax(axis=1), -1).astype(np.int32)
  return act


def replay_on_cpu(workload, global_games, seed, init_step, num_steps):
  """The sampled games stepped on the CPU checker (the reference engine from
  oracle/_ref when it was built, else the C port) with the very policy keys the GPU
  used: returns (obs, mask, cur_player, next_action) after num_steps env-steps."""
  import numpy as np
  from oracle import oracle as O
  kind = "ref" if O.have_ref() else "oracle"
  g = np.asarray(global_games, np.int64)
  b = O.CpuBatch(kind, WORKLOADS[workload], len(g), (1 + g).astype(np.int32))
  b.reset()
  act = policy_draw(b.legal_mask(), seed, init_step, g)
  for t in range(num_steps):
    _, _, _, illegal = b.step(act)
    if illegal.any():
      raise AssertionError("CPU replay: illegal action at step %d" % t)
    act = policy_draw(b.legal_mask(), seed, t, g)
  return b.encode(), b.legal_mask(), b.cur_player(), act, kind


# -------------------------------------------------------------------- GPU arm
def main():
  args = parse_args()
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  if args.cpu_python_worker:
    mode, seconds, seed = args.cpu_python_worker.split(":")
    python_worker(mode, float(seconds), int(seed), args.workload)
    return
  if args.impl == "reference":
    run_reference_arm(args, rank)
    return

  import copy
  import torch

  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device; the engine has no CPU fallback")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  dist = None
  if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
    dist.init_process_group("nccl", device_id=dev)
  cores = None
  if world > 1:
    from hanabi_b200 import sharding
    cores = sharding.bind_rank_to_local_cores(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
  ctx = dict(rank=rank, local_rank=local_rank, world=world, dev=dev, dist=dist,
             cores=sorted(cores) if cores else None)

  line = gpu_arm(args, ctx, extra=False)
  if args.policy == "random" and args.workload == "hanabi_full_2p" and not args.no_policy_extras:
    # BASELINE.json configs[3]: a short Rainbow-policy rollout (bf16 library GEMMs, our step /
    # select kernels, one CUDA graph per step) on every rank, reported beside the headline
    a2 = copy.copy(args)
    a2.policy, a2.mlp_dtype, a2.games = "rainbow", "bf16", min(args.games, 1 << 18)
    a2.steps, a2.warmup, a2.record = min(args.steps, 100), min(max(args.warmup, 3), 10), False
    torch.cuda.empty_cache()
    extra_line = gpu_arm(a2, ctx, extra=True)
    if line is not None and extra_line is not None:
      line["policies"] = {"rainbow": {k: extra_line[k] for k in
                                      ("value", "unit", "ms_per_step", "steps", "warmup", "gpu_launches")}}
      line["policies"]["rainbow"].update(
          games_per_gpu=a2.games, mlp_dtype="bf16", policy=extra_line["config"]["policy"],
          step_kernel_ms=extra_line["roofline"]["kernel_ms"],
          step_kernel_frac=extra_line["roofline"]["frac"])
  if line is not None:
    print(json.dumps(line), flush=True)
  if dist:
    dist.destroy_process_group()


def gpu_arm(args, ctx, extra):
  """One measured configuration on this rank's GPU; returns the JSON line on rank 0
  (None elsewhere).  extra=True: device-timed part only (no e2e, CPU legs, parity)."""
  import numpy as np
  import torch
  from hanabi_b200 import pyhanabi as ph
  rank, local_rank, world, dev, dist = (ctx[k] for k in ("rank", "local_rank", "world", "dev", "dist"))

  # ad-hoc team: below 32 768 games per GPU the host sync of the row gather costs more than
  # running the MLP on every game
  adhoc_all_rows = args.adhoc_all_rows or args.games < 32768
  use_graph = (not args.no_graph) and (args.policy == "rainbow" or
                                       (args.policy == "adhoc" and adhoc_all_rows))
  # (the ppo policy's categorical draw takes the step number as its random key: host launches)
  if use_graph:
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))  # capture needs a non-default stream
  params = WORKLOADS[args.workload]
  n = args.games
  from hanabi_b200 import sharding
  seeds = np.asarray(sharding.shard_seeds(1, world * n, rank, world), np.int32)  # 1 + global index
  fused_policy = args.policy == "random" and not args.separate_policy
  S = max(1, args.shards) if fused_policy else 1
  while n % (S * 32):
    S -= 1
  stats = torch.zeros(32, dtype=torch.int64, device=dev)

  def make_shard(lo, hi):
    m = hi - lo
    b = ph.HanabiBatch(params, m, seeds[lo:hi], device=local_rank)
    b.set_game_index_base(rank * n + lo)  # the policy key's game = global index
    t = dict(obs=torch.empty((m, b.obs_size), dtype=torch.uint8, device=dev),
             mask=torch.empty((m, b.max_moves), dtype=torch.uint8, device=dev),
             cur=torch.empty(m, dtype=torch.int32, device=dev),
             rew=torch.empty(m, dtype=torch.int32, device=dev),
             done=torch.empty(m, dtype=torch.uint8, device=dev),
             act=torch.empty(m, dtype=torch.int32, device=dev))
    b.set_stats_buffer(stats)
    b.reset()
    b.observe(t["obs"], t["mask"], t["cur"])
    return b, t

  batch, T0 = make_shard(0, n // S)
  obs, mask, cur, rew, done, act = (T0[k] for k in ("obs", "mask", "cur", "rew", "done", "act"))
  D, A = batch.obs_size, batch.max_moves

  # pre-bound launches: pointer casts and argument checks happen once, so the
  # Python cost per step stays far below the ~0.25 ms the GPU needs
  bound_step = batch.bind_step_encode(act, rew, done, obs, mask, cur)
  step_encode = lambda t: bound_step()
  launches_per_step = 2
  recorder = None
  if fused_policy:
    # the uniform-random policy rides in the step kernel: it applies act[i] and leaves
    # the next random legal move in act[i] (same draw as BatchSelectAction, epsilon >= 1).
    # S independent shards, each on its own stream: consecutive steps of one shard are
    # ordered by its stream, different shards overlap freely.
    shard_list = [(batch, T0)] + [make_shard(i * (n // S), (i + 1) * (n // S)) for i in range(1, S)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)] if S > 1 else [torch.cuda.current_stream()]
    runs = []
    for (b_i, t_i), st in zip(shard_list, streams):
      b_i.random_legal_actions(t_i["act"], POLICY_SEED, INIT_POLICY_STEP)
      runs.append(b_i.bind_step_encode_random(t_i["act"], POLICY_SEED, t_i["rew"], t_i["done"], t_i["obs"],
                                              t_i["mask"], t_i["cur"], stream=st))
    torch.cuda.synchronize()

    def step_encode(t):
      for r in runs:
        r(t)
    select = lambda t: None
    launches_per_step = S
  elif args.policy == "random":
    select = ph.bind_select_action(None, mask, 1.0, seed=7, out=act)
  elif args.policy == "ppo":
    # the PPO / REINFORCE collect policy (training/hanabi_small/train_ppo_agent_small_env.py:116,
    # tf_agents_lib/masked_networks.py): actor MLP obs -> 150 -> 75 -> A with fixed random weights
    # (layer widths zero-padded to multiples of 32 for the library GEMMs), then the masked
    # categorical draw + log-probability in BatchSampleMaskedCategorical
    torch.manual_seed(0)
    mlp = args.mlp_dtype
    tdt = {"fp32": torch.float32, "tf32": torch.float32, "bf16": torch.bfloat16,
           "fp16": torch.float16}[mlp]
    torch.backends.cuda.matmul.allow_tf32 = mlp == "tf32"
    pad = lambda v: (v + 31) // 32 * 32
    Kp, H1, H2, Ap = pad(D), pad(150), pad(75), pad(A)

    def dense(i, o, ip, op):
      w = torch.zeros(ip, op, device=dev)
      w[:i, :o] = torch.randn(i, o, device=dev) / i ** 0.5
      return w.to(tdt), torch.zeros(op, device=dev, dtype=tdt)
    (w1, c1), (w2, c2), (w3, c3) = dense(D, 150, Kp, H1), dense(150, 75, H1, H2), dense(75, A, H2, Ap)
    x = torch.empty((n, Kp), dtype=tdt, device=dev)
    logits = torch.empty((n, Ap), dtype=tdt, device=dev)
    logp = torch.empty(n, dtype=torch.float32, device=dev)
    batch.observe_net(x)
    step_encode = lambda t: batch.step_encode_net(act, x, rew, done, mask=mask, cur_player=cur)
    launches_per_step = 2  # ours: step (+ network input), masked categorical (+ 3 library GEMMs)

    def select(t):
      h1 = torch._addmm_activation(c1, x, w1)
      h2 = torch._addmm_activation(c2, h1, w2)
      torch.addmm(c3, h2, w3, out=logits)
      ph.sample_masked_categorical(logits, mask, seed=7, step=t, actions=act, log_probs=logp)
  else:
    # Rainbow actor (agents/rainbow_copy/rainbow_agent.py:36-68,163-172): MLP
    # obs_dim -> 512 -> A x 51 with fixed random weights.  The step kernel writes the first
    # GEMM's input (0/1 in the GEMM's element type, K zero-padded to a multiple of 32), the two
    # GEMMs are library calls (cuBLASLt through torch; bias + ReLU in the first one's
    # epilogue), and BatchSelectActionC51 reads the N-padded logits in place:
    # softmax . support, legal mask, argmax.
    torch.manual_seed(0)
    mlp = args.mlp_dtype
    tdt = {"fp32": torch.float32, "tf32": torch.float32, "bf16": torch.bfloat16,
           "fp16": torch.float16}[mlp]
    torch.backends.cuda.matmul.allow_tf32 = mlp == "tf32"
    atoms = 51
    Kp, Np = (D + 31) // 32 * 32, (A * atoms + 63) // 64 * 64
    w1 = torch.zeros(Kp, 512, device=dev)
    w1[:D] = torch.randn(D, 512, device=dev) / D ** 0.5
    w2 = torch.zeros(512, Np, device=dev)
    w2[:, :A * atoms] = torch.randn(512, A * atoms, device=dev) / 512 ** 0.5
    w1, w2 = w1.to(tdt), w2.to(tdt)
    b1 = torch.zeros(512, device=dev, dtype=tdt)
    b2 = torch.zeros(Np, device=dev, dtype=tdt)
    support = torch.linspace(-25.0, 25.0, atoms, device=dev)
    bits = torch.empty((n, batch.packed_words), dtype=torch.int32, device=dev)
    x = torch.empty((n, Kp), dtype=tdt, device=dev)
    logits = torch.empty((n, Np), dtype=tdt, device=dev)
    # the step kernel writes the first GEMM's input itself (BatchStepEncodeNet); the packed
    # rows ride along for the recorder
    want_bits = args.record or args.policy == "adhoc"
    batch.observe_net(x, obs_bits=bits if want_bits else None)
    greedy = ph.bind_select_action_c51(logits, atoms, mask, 0.0, seed=7, out=act, support=support)
    if args.policy == "adhoc" and not adhoc_all_rows:
      # the gathered-rows mode expands only the Rainbow seat's rows: packed rows suffice
      step_encode = lambda t: batch.step_encode_packed(act, rew, done, bits, mask, cur)
    else:
      step_encode = lambda t: batch.step_encode_net(act, x, rew, done,
                                                    obs_bits=bits if want_bits else None,
                                                    mask=mask, cur_player=cur)
    launches_per_step = 2  # ours: step (+ network input), select (+ 2 library GEMMs)

    def rainbow(t):
      h = torch._addmm_activation(b1, x, w1)  # relu(x @ w1 + b1), one cuBLASLt call
      torch.addmm(b2, h, w2, out=logits)
      greedy(t)

    select = rainbow
    if args.record and args.policy == "rainbow":
      # the complete actor loop: policy -> record (o, l, a) of the acting seat -> env step ->
      # rewards to every seat, finished episodes into the device replay memory
      from hanabi_b200 import actor, replay as replay_mod
      memory = replay_mod.DeviceReplayMemory(A, D, args.replay_capacity, update_horizon=1, gamma=0.99,
                                             device=local_rank)
      recorder = actor.DeviceRecorder(n, batch.num_players, D, A, memory, device=local_rank)
      launches_per_step = 18  # + record, after, 3 scan passes, owner, post, leaves, 6 + 1 tree levels, commit

      def select(t):
        rainbow(t)
        recorder.before_step(cur, bits, mask, act)

    if args.policy == "adhoc":
      # ad-hoc team 1: seat 0 = Rainbow, every other seat = RuleBasedAgent, one shared
      # object (training/run_adhoc_experiment.py:195-258); the rule-based kernel
      # overwrites the action of the games whose player to act is one of its seats
      batch.enable_history()
      batch.reset()
      batch.observe(obs, mask, cur)
      batch.observe_net(x, obs_bits=bits)
      agent_state = torch.zeros((n, 1), dtype=torch.int32, device=dev)
      seat_mask = sum(1 << s for s in range(1, batch.num_players))
      launches_per_step = 3

      def select_all(t):
        rainbow(t)
        batch.agent_act(0, seat_mask, act, agent_state=agent_state)

      def select_compact(t):
        # only the games whose player to act is seat 0 go through the network: gather
        # their packed rows and masks, run the MLP on that many rows, scatter the actions
        idx = (cur == 0).nonzero().squeeze(1)  # one host sync per step (the row count)
        m = int(idx.numel())
        if m:
          # row count rounded up to a multiple of 2048: few distinct GEMM shapes, so the
          # library's per-shape kernel selection is paid once each
          mp = (m + 2047) // 2048 * 2048
          xb = torch.zeros((mp, Kp), dtype=tdt, device=dev)
          ph.expand_packed(bits.index_select(0, idx), D, xb[:m])
          hh = torch._addmm_activation(b1, xb, w1)
          lg = torch.addmm(b2, hh, w2)
          am = ph.select_action_c51(lg[:m], atoms, mask.index_select(0, idx), 0.0, seed=7, step=t,
                                    support=support)
          act.index_copy_(0, idx, am)
        batch.agent_act(0, seat_mask, act, agent_state=agent_state)

      select = select_all if adhoc_all_rows else select_compact

  posted_total = [0]

  def after_step():
    # rewards to the seats, finished episodes into the replay: asynchronous, device-driven
    if recorder is not None:
      recorder.after_step(rew, done, return_count=False)

  def one_step(t):
    select(t)
    step_encode(t)
    after_step()

  W = max(args.warmup, 3)
  K = args.steps
  for t in range(W):
    one_step(t)
  stats.zero_()
  posted_total[0] = 0
  torch.cuda.synchronize()

  graph_kernel_ms = None
  if use_graph:
    # the step kernel's own duration from a short evented pass, then one step captured
    pre_ev = []
    for t in range(20):
      select(W + t)
      e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e_a.record(); step_encode(W + t); e_b.record()
      after_step()
      pre_ev.append((e_a, e_b))
    torch.cuda.synchronize()
    graph_kernel_ms = sum(x.elapsed_time(y) for x, y in pre_ev) / len(pre_ev)
    step_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(step_graph, stream=torch.cuda.current_stream()):
      one_step(0)
    # several steps per graph where the step is launch-bound (nothing in a captured step depends
    # on the step number: greedy select, stateless agents)
    graph_steps = args.graph_steps if args.graph_steps > 0 else (8 if n < 32768 else 1)
    multi_graph = None
    if graph_steps > 1:
      multi_graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(multi_graph, stream=torch.cuda.current_stream()):
        for _ in range(graph_steps):
          one_step(0)
      multi_graph.replay()
    for _ in range(3):
      step_graph.replay()
    stats.zero_()
    posted_total[0] = 0
    torch.cuda.synchronize()

  # Everything the timed region needs exists before the ranks line up: the NVML sampler, the
  # events, and NCCL's kernels for the two collectives used below (a first use would load them).
  sampler = ClockSampler(local_rank)
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev2 = torch.cuda.Event(enable_timing=True)
  k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(0 if (fused_policy or use_graph) else K)]
  main_stream = torch.cuda.current_stream()
  line_up = torch.zeros(1, dtype=torch.int64, device=dev)
  if dist:
    warm_stats = torch.zeros_like(stats)
    sharding.reduce_episode_statistics(warm_stats)  # int64 sum all-reduce, as after the steps
    warm_max = torch.zeros(1, dtype=torch.float64, device=dev)
    dist.all_reduce(warm_max, op=dist.ReduceOp.MAX)
    dist.all_reduce(line_up)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
  sampler.start()
  if dist:
    # the window opens in-stream behind one more tiny all-reduce: every rank's ev0 fires when
    # that collective completes on its GPU, no host synchronisation in between
    dist.all_reduce(line_up)
  ev0.record()
  if fused_policy:
    # one kind of launch in the timed region: its average duration is elapsed / launches,
    # no per-launch events needed (they would only put gaps between the kernels)
    if S > 1:
      for st in streams:
        st.wait_event(ev0)
    for t in range(K):
      step_encode(W + t)
    if S > 1:
      for st in streams:
        main_stream.wait_stream(st)
  elif use_graph:
    if multi_graph is not None:
      for g_i in range(K // graph_steps):
        multi_graph.replay()
      for g_i in range(K % graph_steps):
        step_graph.replay()
    else:
      for g_i in range(K):
        step_graph.replay()
  else:
    for t in range(K):
      select(W + t)
      k_ev[t][0].record()
      step_encode(W + t)
      k_ev[t][1].record()
      after_step()
  ev1.record()
  # the only collective, off the step path: 256 B of episode statistics
  # over NCCL (no-op at N=1), timed on its own
  sharding.reduce_episode_statistics(stats)
  ev2.record()
  torch.cuda.synchronize()
  clocks = sampler.stop()
  elapsed_ms = ev0.elapsed_time(ev1)
  stats_allreduce_us = ev1.elapsed_time(ev2) * 1e3 if dist else 0.0
  own_ms = elapsed_ms
  # fused policy: S concurrent launches cover one step of all games, so the step time is
  # the launch time of the kernel over n games
  if fused_policy:
    kernel_ms = elapsed_ms / K
  elif use_graph:
    kernel_ms = graph_kernel_ms
  else:
    kernel_ms = sum(a.elapsed_time(b) for a, b in k_ev) / K
  rank_ms = [elapsed_ms]
  if dist:
    tmax = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    gathered = [torch.zeros_like(tmax) for _ in range(world)]
    dist.all_gather(gathered, tmax)
    rank_ms = [float(x.item()) for x in gathered]
    elapsed_ms = max(rank_ms)  # the job is as fast as its slowest rank
  value = world * n * K / (elapsed_ms * 1e-3)
  if fused_policy:
    kernel_ms = elapsed_ms / K  # the slowest rank's, like the value
  st = stats.cpu().tolist()  # episode statistics of the timed region, summed over the ranks

  # ---- parity of sampled games: every rank replays some of ITS games (global indices,
  # global policy keys) on the CPU reference and compares what the GPU holds after the
  # W + K steps above -- a sharded run must equal the unsharded one game for game.
  parity = None
  if fused_policy and not extra and not args.no_parity_check:
    torch.cuda.synchronize()
    total_steps = W + K
    m = n // S
    per_shard = max(1, args.parity_games // S)
    ok, checked, kind = True, 0, None
    for si, (b_i, t_i) in enumerate(shard_list):
      local = np.unique(np.concatenate([np.array([0, m - 1]),
                                        np.linspace(0, m - 1, per_shard).astype(np.int64)]))
      glob = rank * n + si * m + local
      c_obs, c_mask, c_cur, c_act, kind = replay_on_cpu(args.workload, glob, POLICY_SEED,
                                                        INIT_POLICY_STEP, total_steps)
      idx = torch.as_tensor(local, device=dev)
      ok = ok and bool((t_i["obs"][idx].cpu().numpy() == c_obs).all())
      ok = ok and bool((t_i["mask"][idx].cpu().numpy() == c_mask).all())
      ok = ok and bool((t_i["cur"][idx].cpu().numpy() == c_cur).all())
      ok = ok and bool((t_i["act"][idx].cpu().numpy() == c_act).all())
      checked += len(local)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
    if dist:
      dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    parity = {"ok": bool(flag.item()), "games_per_rank": checked, "steps_replayed": total_steps,
              "checker": "reference engine (oracle/_ref)" if kind == "ref" else "C port (oracle/)",
              "compared": "observation bytes, legal mask, player to act, next random action",
              "ranks": world}
    if not parity["ok"]:
      raise SystemExit("bench.py: rank %d: sampled games differ from the CPU reference" % rank)
  # ---- e2e: the same step through the C-ABI with HOST buffers (pinned); the
  # timed region holds the host policy, the action upload, the kernel and the
  # download of observation, mask, reward, done and player index.
  ne = min(args.e2e_games, n) if args.e2e_games > 0 else n
  e2e = None
  if args.e2e_steps > 0 and not extra:
    hb = ph.HanabiBatch(params, ne, seeds[:ne], device=local_rank)
    hb.set_game_index_base(rank * n)
    pin = lambda *shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    h_obs, h_mask = pin(ne, D, dt=torch.uint8), pin(ne, A, dt=torch.uint8)
    h_cur, h_rew = pin(ne, dt=torch.int32), pin(ne, dt=torch.int32)
    h_done, h_act = pin(ne, dt=torch.uint8), pin(ne, dt=torch.int32)
    hb.reset_host(h_obs, h_mask, h_cur)
    t_policy = [0.0]
    g0 = rank * n  # global index of this rank's first game: the host policy's key
    whole_policy = ph.bind_host_random_legal_actions(h_mask, POLICY_SEED, h_act, first_game=g0)

    def host_step(t):
      # uniform random legal move per game, chosen on the host from the host mask
      p0 = time.perf_counter()
      whole_policy(t)
      t_policy[0] += time.perf_counter() - p0
      hb.step_encode_host(h_act, h_rew, h_done, h_obs, h_mask, h_cur)

    for t in range(3):
      host_step(t)
    if dist:
      dist.barrier()
    t_policy[0] = 0.0
    t0 = time.perf_counter()
    for t in range(args.e2e_steps):
      host_step(3 + t)
    dt = time.perf_counter() - t0
    if dist:
      tm = torch.tensor([dt], dtype=torch.float64, device=dev)
      dist.all_reduce(tm, op=dist.ReduceOp.MAX)
      dt = float(tm.item())
    single_call = {"value": world * ne * args.e2e_steps / dt, "ms_per_step": dt / args.e2e_steps * 1e3,
                   "host_policy_ms_per_step": t_policy[0] / args.e2e_steps * 1e3,
                   "note": "one synchronous BatchStepEncodeHost call per step, host policy before it"}
    # pipelined form of the same step (BatchStepEncodeHostRange + BatchHostWait): the games go
    # through in ranges; the host policy of range c runs while the transfers of range c-1 are
    # on the link.  Every step still uploads every action and delivers every observation,
    # mask, reward, done flag and player index into the host buffers.  Measured with both
    # forms of the link (BatchSetHostLink): packed = the library's default, the observation
    # rows cross PCIe bit-packed and the host worker pool expands them into h_obs inside
    # BatchHostWait; bytes = the kernel's uint8 rows cross the link as they are.
    launch, wait = hb.bind_step_encode_host_range(h_act, h_rew, h_done, h_obs, h_mask, h_cur)
    n_ranges = 16 if ne >= (1 << 19) else 8
    rs = (ne // n_ranges + 127) // 128 * 128
    ranges = [(lo, min(lo + rs, ne) - lo) for lo in range(0, ne, rs)]
    policies = [ph.bind_host_random_legal_actions(h_mask[lo:lo + cnt], POLICY_SEED, h_act[lo:lo + cnt],
                                                  first_game=g0 + lo) for lo, cnt in ranges]

    def piped_step(t):
      for (lo, cnt), pol in zip(ranges, policies):
        wait(lo)  # the previous step's results of this range are in host memory
        pol(t)
        launch(lo, cnt)

    def timed_piped(packed_link, steps, t_base):
      wait(-1)
      hb.set_host_link(packed_link)
      for t in range(3):
        piped_step(t_base + t)
      wait(-1)
      if dist:
        dist.barrier()
      t0 = time.perf_counter()
      for t in range(steps):
        piped_step(t_base + 3 + t)
      wait(-1)
      own = time.perf_counter() - t0
      worst = own
      if dist:
        tm = torch.tensor([own], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        worst = float(tm.item())
      return own, worst

    e2e_steps = 3 * args.e2e_steps
    own_b, dt_b = timed_piped(False, e2e_steps, 100)
    own_dt, dt = timed_piped(True, e2e_steps, 1000)
    Wp = hb.packed_words
    d2h_bytes = ne * (D + A + 4 + 1 + 4)          # bytes delivered to the host buffers / on the link as bytes
    d2h_packed = ne * (4 * Wp + A + 4 + 1 + 4)    # on the link with packed rows
    # what one large pinned copy reaches on this box, for comparison
    probe_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    probe_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    best = 0.0
    for _ in range(3):
      c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      c0.record(); probe_h.copy_(probe_d, non_blocking=True); c1.record()
      torch.cuda.synchronize()
      best = max(best, probe_d.numel() / (c0.elapsed_time(c1) * 1e-3) / 1e9)
    del probe_d, probe_h
    bytes_link = {"value": world * ne * e2e_steps / dt_b, "unit": "env-steps/s",
                  "ms_per_step": dt_b / e2e_steps * 1e3, "d2h_bytes_per_step": d2h_bytes,
                  "d2h_gbs_achieved": d2h_bytes * e2e_steps / own_b / 1e9,
                  "note": "BatchSetHostLink(0): the kernel's uint8 rows cross the link; bound by the "
                          "observation download"}
    e2e = {"value": world * ne * e2e_steps / dt, "unit": "env-steps/s",
           "h2d_bytes_per_step": ne * 4, "d2h_bytes_per_step": d2h_packed,
           "host_bytes_delivered_per_step": d2h_bytes,
           "games_per_step": ne, "steps": e2e_steps,
           "ms_per_step": dt / e2e_steps * 1e3,
           "link": "packed (library default): %d B of bit-packed row per game on PCIe instead of %d B, "
                   "expanded into the caller's uint8 rows by the host worker pool inside BatchHostWait"
                   % (4 * Wp, D),
           "pcie": {"d2h_gbs_achieved": d2h_packed * e2e_steps / own_dt / 1e9,
                    "d2h_gbs_one_large_memcpy": best,
                    "host_rows_gbs": ne * D * e2e_steps / own_dt / 1e9,
                    "note": "this rank's link; with packed rows the step is bound by the host's "
                            "expansion + policy threads, not by the link"},
           "host_policy_threads": ph.host_policy_threads(),
           "host_cores_of_rank0": ctx.get("cores"),
           "note": "BatchStepEncodeHostRange/BatchHostWait with pinned host buffers, %d ranges per "
                   "step; host random-legal policy (persistent worker pool) inside the timed region, "
                   "overlapping the transfers of the previous range; uint8 observation rows, legal "
                   "mask, reward, done, player index of every game delivered to host memory every "
                   "step" % len(ranges),
           "bytes_link": bytes_link,
           "single_call": single_call}
    del hb

  # ---- optional extra, reported separately and never mixed into the headline:
  # bit-packed observation rows (16-byte multiples, 96 B instead of 658 B for 2p)
  packed = None
  if args.packed and not extra:
    Wp = batch.packed_words
    n_p = batch.num_games  # the first shard when the games are split
    bits = torch.empty((n_p, Wp), dtype=torch.int32, device=dev)
    step_packed = lambda: batch.step_encode_packed(act, rew, done, bits, mask, cur)
    if fused_policy:
      select = ph.bind_select_action(None, mask, 1.0, seed=7, out=act)
    for t in range(W):
      select(10**6 + t); step_packed()
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K_pk = min(K, 500)
    p0.record()
    for t in range(K_pk):
      select(10**6 + W + t); step_packed()
    p1.record()
    torch.cuda.synchronize()
    pms = p0.elapsed_time(p1) / K_pk
    packed = {"value": world * n_p / (pms * 1e-3), "unit": "env-steps/s", "ms_per_step": pms,
              "games": n_p,
              "obs_bytes_per_game": 4 * Wp, "note": "device-resident, bit-packed observation rows"}
    if args.e2e_steps > 0:
      hb2 = ph.HanabiBatch(params, ne, seeds[:ne], device=local_rank)
      hb2.set_game_index_base(rank * n)
      pinned = lambda *shape, dt: torch.empty(shape, dtype=dt).pin_memory()
      q_bits, q_mask = pinned(ne, Wp, dt=torch.int32), pinned(ne, A, dt=torch.uint8)
      q_cur, q_rew = pinned(ne, dt=torch.int32), pinned(ne, dt=torch.int32)
      q_done, q_act = pinned(ne, dt=torch.uint8), pinned(ne, dt=torch.int32)
      hb2.reset_host(None, q_mask, q_cur)
      # pipelined (BatchStepEncodeHostRangePacked + BatchHostWait), as the byte-row e2e above
      launch_p, wait_p = hb2.bind_step_encode_host_range(q_act, q_rew, q_done, None, q_mask, q_cur,
                                                         obs_bits=q_bits)
      rs_p = (ne // 8 + 127) // 128 * 128
      ranges_p = [(lo, min(lo + rs_p, ne) - lo) for lo in range(0, ne, rs_p)]
      pols_p = [ph.bind_host_random_legal_actions(q_mask[lo:lo + cnt], POLICY_SEED, q_act[lo:lo + cnt],
                                                  first_game=rank * n + lo) for lo, cnt in ranges_p]

      def piped_packed(t):
        for (lo, cnt), pol in zip(ranges_p, pols_p):
          wait_p(lo)
          pol(t)
          launch_p(lo, cnt)

      for t in range(3):
        piped_packed(200 + t)
      wait_p(-1)
      if dist:
        dist.barrier()
      steps_p = 3 * args.e2e_steps
      t0 = time.perf_counter()
      for t in range(steps_p):
        piped_packed(203 + t)
      wait_p(-1)
      dtp = time.perf_counter() - t0
      own_dtp = dtp
      if dist:
        tm = torch.tensor([dtp], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dtp = float(tm.item())
      d2h_p = ne * (4 * Wp + A + 9)
      packed["e2e"] = {"value": world * ne * steps_p / dtp, "unit": "env-steps/s",
                       "d2h_bytes_per_step": d2h_p, "h2d_bytes_per_step": ne * 4,
                       "games_per_step": ne, "ms_per_step": dtp / steps_p * 1e3,
                       "d2h_gbs_achieved": d2h_p * steps_p / own_dtp / 1e9,
                       "note": "BatchStepEncodeHostRangePacked/BatchHostWait, %d ranges per step; the "
                               "host-buffer form for consumers that unpack bits themselves" % len(ranges_p)}
      del hb2

  if rank != 0:
    return None

  peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(peaks_path):
    peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
  else:
    peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
  algo = ALGO_BYTES[args.workload]
  algo_note = "uint8 observation rows"
  if args.policy != "random":
    # the step kernel of these modes writes other observation rows than one byte per bit
    esz = {"fp32": 4, "tf32": 4, "bf16": 2, "fp16": 2}[args.mlp_dtype]
    packed_bytes = 4 * batch.packed_words
    if args.policy == "adhoc" and not adhoc_all_rows:
      algo, algo_note = algo - D + packed_bytes, "bit-packed observation rows"
    else:
      with_bits = bool(args.record or args.policy == "adhoc")
      algo = algo - D + Kp * esz + (packed_bytes if with_bits else 0)
      algo_note = "network-input rows (%d x %d B)%s" % (Kp, esz, " + bit-packed rows" if with_bits else "")
  achieved = algo * n / (kernel_ms * 1e-3) / 1e9
  traffic = None
  tpath = os.path.join(ROOT, "profiles", "traffic.json")
  if os.path.exists(tpath):
    tj = json.load(open(tpath))
    if tj.get("workload") == args.workload and tj.get("games") == n:
      traffic = tj.get("dram_bytes_per_launch")
  line = {
      "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
      "steps": K, "warmup": W, "ms_per_step": elapsed_ms / K, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
      "config": {"workload": args.workload, "games_per_gpu": n, "obs_dim": D, "num_moves": A,
                 "policy": ("uniform random legal, drawn inside the step kernel "
                            "(BatchStepEncodeRandom)" if fused_policy else
                            "uniform random legal (BatchSelectAction, epsilon=1)"
                            if args.policy == "random" else
                            "PPO actor MLP %d->150->75->%d (cuBLASLt %s GEMMs via torch) on the step "
                            "kernel's network-input rows + BatchSampleMaskedCategorical"
                            % (D, A, args.mlp_dtype) if args.policy == "ppo" else
                            "Rainbow C51 MLP %d->512->%dx51 (cuBLASLt %s GEMMs via torch) on the "
                            "step kernel's own network-input rows (BatchStepEncodeNet) + "
                            "BatchSelectActionC51 (epsilon=0)"
                            % (D, A, args.mlp_dtype) +
                            ("; seats 1.. = batched RuleBasedAgent (BatchAgentAct); MLP on %s"
                             % ("all games" if adhoc_all_rows else
                                "the games whose player to act is seat 0 (gathered rows)")
                             if args.policy == "adhoc" else "")),
                 "cuda_graph": bool(use_graph),
                 "graph_steps": (graph_steps if use_graph else None),
                 "auto_reset": True, "seeds": "1 + global game index",
                 "l2": "outputs per step (%.0f MB) exceed the 126 MB L2" % (n * (D + A) / 1e6),
                 "parallelism": "games sharded over %d GPU(s), no step-path collective" % world +
                                ("; per GPU %d independent shards of %d games on their own streams"
                                 % (S, n // S) if S > 1 else "")},
      "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                   "frac": achieved / peak, "traffic": traffic, "kernel": "k_main<step>",
                   "kernel_ms": kernel_ms, "launches_per_step": launches_per_step,
                   "algorithmic_bytes_per_env_step": algo, "observation_output": algo_note,
                   "peak_source": peak_src},
      "e2e": e2e,
      "gpu_launches": launches_per_step * K,
      "clocks": clocks,
      "parity_checked": bool(parity and parity["ok"]),
      "parity": parity,
      "ranks": {"ms_per_step": [x / K for x in rank_ms], "own_window": "CUDA events on each rank's "
                "stream around its K steps, opened in-stream behind a tiny all-reduce; value uses the max",
                "stats_allreduce_us": stats_allreduce_us},
      "episodes": {"count": st[0], "mean_score": st[1] / max(st[0], 1),
                   "mean_length": st[2] / max(st[0], 1)},
  }
  if packed is not None:
    line["packed_obs"] = packed
  if recorder is not None:
    line["actor_loop"] = {"transitions_in_replay": memory.add_count, "replay_capacity": args.replay_capacity,
                          "transitions_dropped": recorder.dropped,
                          "note": "DeviceRecorder + DeviceReplayMemory inside the timed region"}
  if not args.no_cpu_baseline and world == 1 and not extra:  # the CPU legs run at N = 1 only
    line["cpu_baseline"] = cpu_reference_rate(args.workload, args.cpu_seconds)
    if args.python_seconds > 0:
      # the Python paths a reference user runs today (north_star: "pyhanabi stepped from Python,
      # single-process and one process per core")
      line["cpu_baseline"]["python_paths"] = python_baselines(args.workload, args.python_seconds)
  return line


if __name__ == "__main__":
  main()
