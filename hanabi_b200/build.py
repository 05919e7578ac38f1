"""In-tree build of libpyhanabi.so (CUDA kernels + C-ABI) for sm_100a.

The library sits next to this file so that it travels with the repo snapshot and
is found by hanabi_b200.pyhanabi.try_load() the way the reference finds its own
libpyhanabi.so next to pyhanabi.py (hanabi_learning_environment/pyhanabi.py:22-25).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpyhanabi.so")
CUDA_SOURCES = ["hb_batch.cu", "hb_policy.cu", "hb_replay.cu"]
HOST_SOURCES = ["hb_single.cc", "hb_host.cc"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
]


def _nvcc():
  for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError("nvcc not found")


def _sources():
  out = []
  for name in CUDA_SOURCES + HOST_SOURCES:
    path = os.path.join(CSRC, name)
    if os.path.exists(path):
      out.append(path)
  return out


def needs_build():
  if not os.path.exists(LIB):
    return True
  newest = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC))
  return newest > os.path.getmtime(LIB)


def build(force=False, verbose=False, extra_flags=()):
  """Compile every CUDA source into hanabi_b200/libpyhanabi.so."""
  if not force and not needs_build():
    return LIB
  cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-shared", "-o", LIB] + _sources()
  if verbose:
    print(" ".join(cmd), file=sys.stderr)
  subprocess.check_call(cmd)
  return LIB


if __name__ == "__main__":
  build(force="--force" in sys.argv, verbose=True,
        extra_flags=["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else [])
