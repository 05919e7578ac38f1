"""B200-native batched Hanabi engine (hot path of hellovertex/hanabi).

  hanabi_b200.pyhanabi   cffi binding of libpyhanabi.so (CUDA kernels + C-ABI)
  hanabi_b200.rl_env     HanabiEnv-compatible batched environment
  hanabi_b200.build      in-tree nvcc build for sm_100a
"""
__all__ = ["pyhanabi", "rl_env", "build"]
