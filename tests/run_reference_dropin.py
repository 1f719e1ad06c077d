"""Runs the UNMODIFIED reference trainer (baseline/_ref/train_torch.py: RLSystem, train_torch.py:69-675) with this library's
drop-in classes on its own plug points (BASELINE.json configs[4] at n_parallel = 24, one GPU): dropin/ sits ahead of the
reference on sys.path, so utils.get_class("src.mcts", "MCTSSearchVec") (train_torch.py:90), get_class(environment_path,
environment_name) (:93) and `from replay_buffer import ...` (:5) resolve to muzero_breakout_b200, while src.networks,
utils and train_torch itself are the reference's files.  One acting stage (:160-233) + one training stage (:369-452,
which ends with the 2-env test rollout :530-610 on the online network).  Started by tests/test_reference_dropin_gpu.py
in a scratch directory (train_torch creates logs/ in the CWD); prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dropin"), ROOT]
if os.environ.get("MZB_DROPIN_AGENT") == "1":          # the learner-side drop-in too: src.networks -> muzero-breakout_b200/src/agent.py
    sys.path.insert(0, os.path.join(ROOT, "dropin_train"))

import torch  # noqa: E402

from baseline import ref  # noqa: E402

ref_dir = ref.install()                       # appended AFTER dropin/ and the repo root
import train_torch  # noqa: E402  (the reference's file; runs set_seed(42) at import)

cfg = ref.load_cfg()
cfg["num_episodes"] = 1
cfg["num_batches"] = int(os.environ.get("MZB_REF_BATCHES", "2"))
cfg["minibatch_size"] = int(os.environ.get("MZB_REF_MINIBATCH", "64"))
system = train_torch.RLSystem(cfg)
system.max_steps_test = int(os.environ.get("MZB_REF_TEST_STEPS", "12"))     # the reference's own attribute (train_torch.py:83)
mods = dict(mcts=type(system.latent_mcts).__module__, env=type(system.environment).__module__, replay=type(system.replay_buffer).__module__,
            agent=type(system.mu_zero).__module__, trainer=os.path.realpath(train_torch.__file__))

log = dict(searches=0, visit_sum_ok=True, value_finite=True, visits_dtype="", value_device="", steps=0, done_aliased=True, step_shapes_ok=True,
           batches=set())
mcts, env = system.latent_mcts, system.environment
orig_search, orig_step = mcts.search, env.step


def search(hidden_state, mask, it):
    value, visits = orig_search(hidden_state, mask, it)
    log["searches"] += 1
    log["visit_sum_ok"] &= bool((visits.sum(dim=1) == cfg["num_simulations"]).all())
    log["value_finite"] &= bool(torch.isfinite(value).all())
    log["visits_dtype"], log["value_device"] = str(visits.dtype), str(value.device)
    log["batches"].add(int(visits.shape[0]))
    return value, visits


def step(state, action, done_mask):
    out = orig_step(state, action, done_mask)
    log["steps"] += 1
    log["done_aliased"] &= out[2] is done_mask                     # parallel_breakout.py:204,247 mutate and return the caller's mask
    B = action.shape[0]
    log["step_shapes_ok"] &= (tuple(out[0].shape) == (B, 3, 16, 20) and out[0].dtype == torch.float32 and tuple(out[1].shape) == (B,)
                              and out[2].dtype == torch.bool and tuple(out[3].shape) == (B, 3))
    return out


mcts.search, env.step = search, step
graph_step = None
if os.environ.get("MZB_GRAPH_TRAIN") == "1":      # the unmodified _training_stage, each loop iteration ONE CUDA-graph replay
    from muzero_breakout_b200 import train as _train
    if type(system.mu_zero).__module__ == "src.networks":
        _train.accelerate_agent(system.mu_zero)
    graph_step = _train.accelerate_training_stage(system)
t0 = time.perf_counter()
system._acting_stage()
t_act = time.perf_counter() - t0
acting = dict(seconds=t_act, moves=log["steps"], searches=log["searches"], replay_length=int(system.replay_buffer.length),
              trajectories=len(system.observation_trajectories),
              episode_lengths=[int(o.length) for o in system.observation_trajectories][:6])
import muzero_breakout_b200 as _mzb  # noqa: E402
n_launch0 = _mzb.launch_count()
t0 = time.perf_counter()
system._training_stage()
t_train = time.perf_counter() - t0
n_launch_train = _mzb.launch_count() - n_launch0
out = dict(modules=mods, acting=acting, training=dict(seconds=t_train, steps=int(system.training_step), library_launches=int(n_launch_train),
                                                      graph_replays=None if graph_step is None else int(graph_step.replays),
                                                      optimizer_steps=int(getattr(system.mu_zero.optimizer, "step_count", -1))),
           search_calls=log["searches"], env_steps=log["steps"], search_batches=sorted(log["batches"]),
           visit_sum_ok=log["visit_sum_ok"], value_finite=log["value_finite"], visits_dtype=log["visits_dtype"], value_device=log["value_device"],
           done_aliased=log["done_aliased"], step_shapes_ok=log["step_shapes_ok"], env_batch_after=int(system.environment.batch),
           mcts_net_is_target=system.latent_mcts.mu_zero is system.mu_zero_target)
print("RESULT " + json.dumps(out))
