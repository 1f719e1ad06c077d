"""Shadow of the reference's `src/networks.py` for the LEARNER side (optional; put `dropin_train/` ahead of `dropin/` and the reference
on sys.path): `get_class("src.networks", "MuZeroAgent")` (train_torch.py:85) then loads the drop-in agent whose ResidualBlock trunks train
on this library's kernels and whose optimizer is the flat-buffer Adam (muzero-breakout_b200/src/agent.py)."""
from muzero_breakout_b200.src.agent import (ConvBlock, DynamicsNetwork, MuZeroAgent, PredictionNetwork, RepresentationNetwork,  # noqa: F401
                                            ResidualBlock)
