"""Shadow of the reference's `environment/parallel_breakout.py` (config.yaml:53-54, train_torch.py:93-94)."""
from muzero_breakout_b200.environment.parallel_breakout import BreakoutEnvironment, MuZeroEnvironment  # noqa: F401
