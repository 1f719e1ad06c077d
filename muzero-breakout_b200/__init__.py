"""B200-native acting hot path of MuZero-Breakout: vectorised Breakout environment and batched
latent MCTS as hand-written sm_100a CUDA kernels behind a C ABI (include/mzb200.h), with Python
hosts that mirror the reference's plug-in classes:

    muzero_breakout_b200.environment.parallel_breakout.BreakoutEnvironment
        <-> reference environment/parallel_breakout.py:59
    muzero_breakout_b200.src.mcts.MCTSSearchVec
        <-> reference src/mcts.py:10

There is no CPU fallback: importing works anywhere, but every compute call needs a CUDA device and
the built libmzb200.so, and raises otherwise.
"""
from ._lib import build, lib, launch_count  # noqa: F401

__all__ = ["build", "lib", "launch_count"]
