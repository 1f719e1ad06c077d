"""Shadow of the reference's `src/mcts.py`: put `dropin/` BEFORE the reference checkout on sys.path and
`get_class("src.mcts", "MCTSSearchVec")` (train_torch.py:90) resolves to the B200 implementation, while
`src.networks` still comes from the reference (both `src/` directories are namespace-package portions)."""
from muzero_breakout_b200.src.mcts import MCTSSearchVec  # noqa: F401
