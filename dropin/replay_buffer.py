"""Shadow of the reference's top-level `replay_buffer.py` (train_torch.py:5): device-resident trajectory store."""
from muzero_breakout_b200.replay_buffer import ObservationTrajectory, ReplayBuffer  # noqa: F401
