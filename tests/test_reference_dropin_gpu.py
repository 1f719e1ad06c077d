"""BASELINE.json configs[4] at one GPU: the reference's own loop (unmodified train_torch.py RLSystem, :160-257 acting,
:369-452 training + test rollout) running on the drop-in BreakoutEnvironment / MCTSSearchVec / ReplayBuffer.
The reference checkout travels to the GPU box as the git-ignored baseline/_ref (made by __graft_entry__.build())."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("agent_dropin", [False, True], ids=["reference_agent", "dropin_agent"])
def test_unmodified_rlsystem_runs_on_the_dropins(tmp_path, agent_dropin):
    """agent_dropin: additionally `src.networks` -> the learner-side drop-in (dropin_train/): the reference's own _training_stage
    (train_torch.py:369-452: _k_step_rollout, loss_fn, loss.backward(), optimizer.step()) then trains the ResidualBlock trunks on this
    library's kernels and steps the flat-buffer Adam."""
    from baseline import ref
    if ref.ref_dir() is None:
        pytest.skip("no reference checkout on this box (baseline/_ref is made by __graft_entry__.build() in the build container)")
    env = dict(os.environ, PYTHONPATH="", MZB_DROPIN_AGENT="1" if agent_dropin else "0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_dropin.py")], capture_output=True, text=True, cwd=tmp_path, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = next(ln for ln in r.stdout.splitlines() if ln.startswith("RESULT "))
    out = json.loads(line[len("RESULT "):])
    print(json.dumps(out, indent=1))
    m = out["modules"]
    assert m["mcts"] == "muzero_breakout_b200.src.mcts" and m["env"] == "muzero_breakout_b200.environment.parallel_breakout"
    assert m["replay"] == "muzero_breakout_b200.replay_buffer"
    assert m["agent"] == ("muzero_breakout_b200.src.agent" if agent_dropin else "src.networks")
    if agent_dropin:
        assert out["training"]["library_launches"] >= 2 * 5 * 28 * 10, "the trunk kernels did not run inside _training_stage"
    assert os.path.realpath(ref.ref_dir()) in m["trainer"]                     # train_torch.py is the reference's file
    a = out["acting"]
    assert a["moves"] >= 1 and a["searches"] >= a["moves"] and a["trajectories"] == 24
    assert a["replay_length"] > 0, "no trajectory reached the replay buffer"
    assert out["visit_sum_ok"], "root visit counts must sum to num_simulations (train_torch.py:192-198 normalises them)"
    assert out["value_finite"] and out["visits_dtype"] == "torch.int64" and out["value_device"] == "cpu"      # mcts.py:71
    assert out["done_aliased"], "step() must return the caller's done_mask object (train_torch.py:179,201,209)"
    assert out["step_shapes_ok"]
    assert out["training"]["steps"] >= 1
    assert out["search_batches"] == [2, 24]                                     # acting at n_parallel, the test rollout at batch 2 (:448-452)
    assert out["env_batch_after"] == 24 and out["mcts_net_is_target"]


def test_unmodified_training_stage_as_graph_replays(tmp_path):
    """train.accelerate_training_stage: the reference's own, unmodified `_training_stage` loop (train_torch.py:369-452) with every iteration
    ONE CUDA-graph replay (rollout + loss + backward + Adam; the reference's `_k_step_rollout`, `loss_fn` and `optimizer.step()` plug points
    patched on the live objects) -- graph replays == training steps == optimizer updates, then the 2-env test rollout on the updated net."""
    from baseline import ref
    if ref.ref_dir() is None:
        pytest.skip("no reference checkout on this box")
    env = dict(os.environ, PYTHONPATH="", MZB_DROPIN_AGENT="1", MZB_GRAPH_TRAIN="1", MZB_REF_BATCHES="3")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_dropin.py")], capture_output=True, text=True, cwd=tmp_path, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads(next(ln for ln in r.stdout.splitlines() if ln.startswith("RESULT "))[len("RESULT "):])
    t = out["training"]
    assert t["steps"] == 3 and t["graph_replays"] == 3 and t["optimizer_steps"] == 3, t
    assert out["search_batches"] == [2, 24] and out["mcts_net_is_target"] and out["value_finite"]
