"""Zero-change mode: the reference's UNMODIFIED Python stack --
hanabi_learning_environment/pyhanabi.py (cffi ABI mode, text-parsing pyhanabi.h) and
rl_env.py -- loaded against THIS repository's include/pyhanabi.h + libpyhanabi.so, stepped
for 500 moves per configuration and hashed with the survey's protocol (SURVEY.md section 8c,
appendix F: seed 1234, a = legal[(t*7+3) % len(legal)], reset() on done, SHA-256 over
vectorized observation of the player to act, legal mask, <iB(reward, done)).  The digests
recorded from the reference's own library must come out.

Runs where /root/reference exists (the build container); skipped elsewhere.  A subprocess
keeps the reference's module names (`pyhanabi`, `rl_env`) out of this test session.
"""
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_HLE = "/root/reference/hanabi_learning_environment"
SURVEY_HASHES = {
    "full2": ("13e4d72cacea246847f512c55804672532f00f4bb28b040a7e78e902c3033130", 40),
    "full4": ("c4f49d94176c217a802d515f1ba252098023fd3c1dfe52f2f8c70400cb279b21", 25),
    "full5": ("771a59763b9bb64552ce724fa24232421e794f41b102a4f19a7f97b3eb290556", 31),
    "small2": ("792fe9f0b761845a9f307419e38704919c883433ff3f2bcd9c432850d6306116", 186),
}

SCRIPT = r'''
import hashlib, json, struct, sys
ref_hle, dropin = sys.argv[1], sys.argv[2]
sys.path.insert(0, ref_hle)
import pyhanabi                      # the reference's binding, unmodified
assert pyhanabi.try_cdef(prefixes=[dropin]), "cdef failed on this repo's pyhanabi.h"
assert pyhanabi.try_load(prefixes=[dropin]), "this repo's libpyhanabi.so failed to load"
import rl_env                        # the reference's environment, unmodified
CONFIGS = {
    "full2": dict(players=2), "full4": dict(players=4), "full5": dict(players=5),
    "small2": dict(players=2, colors=2, ranks=5, hand_size=2, max_information_tokens=3, max_life_tokens=1),
}
out = {}
for name, cfg in CONFIGS.items():
  env = rl_env.HanabiEnv(dict(cfg, seed=1234))
  A = env.num_moves()
  h = hashlib.sha256()
  def emit(obs, reward, done):
    po = obs["player_observations"][obs["current_player"]]
    mask = bytearray(A)
    for uid in po["legal_moves_as_int"]:
      mask[uid] = 1
    h.update(bytes(bytearray(po["vectorized"])))
    h.update(bytes(mask))
    h.update(struct.pack("<iB", int(reward), int(done)))
    return po["legal_moves_as_int"]
  legal = emit(env.reset(), 0, 0)
  episodes = 0
  for t in range(500):
    obs, reward, done, _ = env.step(int(legal[(t * 7 + 3) % len(legal)]))
    legal = emit(obs, reward, done)
    if done:
      episodes += 1
      legal = emit(env.reset(), 0, 0)
  out[name] = [h.hexdigest(), episodes]
print(json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(REF_HLE), reason="needs the reference tree (build container)")
def test_reference_python_stack_runs_unchanged_on_this_library(built_lib, tmp_path):
  dropin = tmp_path / "dropin"
  dropin.mkdir()
  shutil.copy(os.path.join(ROOT, "include", "pyhanabi.h"), dropin / "pyhanabi.h")
  os.symlink(built_lib, dropin / "libpyhanabi.so")
  env = {k: v for k, v in os.environ.items() if k != "HANABI_B200_LIB"}
  res = subprocess.run([sys.executable, "-c", SCRIPT, REF_HLE, str(dropin)], capture_output=True,
                       text=True, env=env, timeout=600, cwd=str(tmp_path))
  assert res.returncode == 0, res.stderr[-2000:]
  got = json.loads(res.stdout.strip().splitlines()[-1])
  for name, (digest, episodes) in SURVEY_HASHES.items():
    assert got[name] == [digest, episodes], name
