"""GPU tests of the on-device acting loop (muzero-breakout_b200/acting.py, csrc/acting.cu) against the reference's own acting loop
(tests/golden/acting.npz: RLSystem._run_episode with injected search outputs and uniforms), and against a
restatement of the reference's history / recording semantics (train_torch.py:171-233, 259-332) and the CPU
environment oracle."""
import numpy as np
import pytest
import torch

import oracle
from common import perturb_bn
from oracle.networks import OracleAgent
from refshim_rng import rng_u32

pytestmark = pytest.mark.gpu

ENV_CFG = dict(n_parallel=5, paddle_hit_reward=0.0, brick_hit_reward=1.0, game_lost_reward=-1.0, game_won_reward=5.0,
               output_device="cuda")


def make(B, sims, precision, **kw):
    from muzero_breakout_b200.acting import Actor
    from muzero_breakout_b200.environment.parallel_breakout import BreakoutEnvironment
    from muzero_breakout_b200.src.mcts import MCTSSearchVec
    torch.manual_seed(0)
    agent = OracleAgent(); perturb_bn(agent, 1); agent.eval_mode()
    cfg = {"num_simulations": sims, "actions": [0, 1, 2], "latent_resolution": [4, 5], "model": agent.cfg,
           "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985, "precision": precision, "output_device": "cuda"}}
    env = BreakoutEnvironment(dict(ENV_CFG, n_parallel=B))
    return Actor(env, MCTSSearchVec(cfg, agent, None), **kw), agent


def test_episode_history_and_recording_semantics():
    B, sims = 5, 6
    actor, agent = make(B, sims, "f32", temperature=1.0, seed=3, max_moves=45)
    actor.keep_rep_inputs = []
    torch.manual_seed(11)
    out = actor.run_episode()
    T = out["action"].shape[0]
    assert T == len(actor.keep_rep_inputs) and 1 <= T <= 45
    act, rew, recd = out["action"].cpu().numpy(), out["reward"].cpu().numpy(), out["recorded"].cpu().numpy()
    frames = out["frames"].cpu().numpy()
    assert np.all(out["visits"].sum(-1).cpu().numpy() == sims)

    # (1) environment: replay the recorded actions through the CPU oracle from the same initial state
    orc = oracle.EnvOracle(B)
    state = out["initial_state"].cpu().numpy()
    orc.ball_dx[:] = out["initial_dx"].cpu().numpy(); orc.ball_dy[:] = -1.0
    gray0 = oracle.gray(state)
    done = np.zeros(B, np.uint8)
    states_list = [[gray0[b]] * 31 for b in range(B)]          # _pad_initial_state :313-332
    actions_list = [[0] * 32 for _ in range(B)]
    cur = gray0
    for t in range(T):
        # (2) rep-net input = last 31 appended frames + current frame + last 32 actions / 3 (:259-293)
        want = np.stack([np.concatenate([np.concatenate(states_list[b][-31:], 0), cur[b],
                                         np.repeat((np.array(actions_list[b][-32:], np.float32) / np.float32(3.0))[:, None, None], 16, 1).repeat(20, 2)], 0)
                         for b in range(B)])
        got = actor.keep_rep_inputs[t].cpu().numpy()
        live = done == 0
        assert np.array_equal(got[live], want[live]), f"move {t}: rep-net input differs"
        assert np.array_equal(recd[t], live), f"move {t}: recorded mask"
        prev_done = done.copy()
        state, r, done, _ = orc.step(state, act[t], done)
        cur = oracle.gray(state)
        assert np.array_equal(rew[t], r), f"move {t}: reward"
        assert np.array_equal(frames[t], cur), f"move {t}: gray frame"
        for b in range(B):
            if not prev_done[b]:                               # add_observation only for games not yet done (:204-208)
                actions_list[b].append(int(act[t, b])); states_list[b].append(cur[b])
    assert np.array_equal(out["done"].cpu().numpy().astype(np.uint8), done)
    assert T == 45 or done.all()                                # while not all(done) and length <= max (:184-187)


@pytest.mark.parametrize("temperature", [1.0, 0.5, 0.25])
def test_sample_actions_matches_restatement(temperature):
    from muzero_breakout_b200 import _lib
    B, seed, step = 50000, 77, 5
    g = torch.Generator().manual_seed(1)
    cuts = torch.sort(torch.randint(0, 51, (B, 2), generator=g), dim=1)[0]
    visits = torch.stack([cuts[:, 0], cuts[:, 1] - cuts[:, 0], 50 - cuts[:, 1]], 1).to(torch.int64)
    action = torch.empty(B, dtype=torch.int64, device="cuda")
    probs = torch.empty((B, 3), dtype=torch.float32, device="cuda")
    slot = torch.empty(B, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().mz_sample_actions(B, visits.cuda().data_ptr(), temperature, seed, step, action.data_ptr(), slot.data_ptr(),
                                            probs.data_ptr(), torch.cuda.current_stream().cuda_stream))
    a = action.cpu().numpy()
    assert np.array_equal(a, slot.cpu().numpy())
    w = visits.float() ** (1 / temperature)                     # train_torch.py:192-193
    p = (w / w.sum(1, keepdim=True)).numpy()
    assert np.allclose(probs.cpu().numpy(), p, atol=2e-6)
    u = np.array([(rng_u32(seed, b, step) >> 8) / 16777216.0 for b in range(B)], np.float32)
    want = np.where(u < p[:, 0], 0, np.where(u < p[:, 0] + p[:, 1], 1, 2))
    near = (np.abs(u - p[:, 0]) < 1e-5) | (np.abs(u - p[:, 0] - p[:, 1]) < 1e-5)     # pow() last-bit differences
    assert np.array_equal(a[~near], want[~near])
    assert np.all(visits.numpy()[np.arange(B), a] > 0)           # never an unvisited action
    # frequencies follow the probabilities
    for k in range(3):
        assert abs((a == k).mean() - p[:, k].mean()) < 0.01


def test_bf16_episode_runs_and_is_consistent():
    B, sims = 64, 8
    actor, _ = make(B, sims, "bf16", temperature=0.5, seed=1, max_moves=30, check_done_every=4, record_frames=False)
    out = actor.run_episode()
    T = out["action"].shape[0]
    assert out["visits"].shape == (T, B, 3) and "frames" not in out
    rec = out["recorded"].cpu()
    assert rec[0].all() and torch.all(rec[:-1] | ~rec[1:])      # once done, never recorded again
    assert torch.all(out["visits"].sum(-1) == sims)
    out2 = actor.run_episode()                                   # a second episode reuses the buffers
    assert out2["action"].shape[1] == B


# ------------------------------------------------------------------------------------------------------------------------------------
# pinned against the reference's own acting loop: tests/golden/acting.npz = one episode of RLSystem._run_episode (train_torch.py:171-233)
# on the reference environment with injected search outputs and an injected uniform stream (tests/golden/gen_golden.py acting)
class _PresetSearch:
    """stands in for MCTSSearchVec in Actor: the golden's per-move (value, visit counts) instead of a search"""

    def __init__(self, real, value, visits):
        self.real, self.value, self.visits, self.t = real, value, visits, 0

    def packed_networks(self):
        return self.real.packed_networks()

    def search(self, hidden, mask, it):
        t = self.t
        self.t += 1
        return torch.from_numpy(self.value[t]).cuda(), torch.from_numpy(self.visits[t]).cuda()


def test_actor_episode_matches_reference_run_episode(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "acting.npz"))
    B, T, seed, env_seed, sims = (int(x) for x in g["meta"])
    actor, _ = make(B, sims, "f32", temperature=float(g["temperature"]), seed=seed, max_moves=261)
    actor.mcts = _PresetSearch(actor.mcts, g["value"], g["visits"])
    actor.keep_rep_inputs = []
    torch.manual_seed(env_seed)                                  # the reference's reset() draws (parallel_breakout.py:116-136)
    out = actor.run_episode()
    assert np.array_equal(out["initial_state"].cpu().numpy(), g["initial_state"])
    assert out["action"].shape[0] == T, "the episode ends when every game is done (train_torch.py:184)"
    act, recd = out["action"].cpu().numpy(), out["recorded"].cpu().numpy()
    rew, frames = out["reward"].cpu().numpy(), out["frames"].cpu().numpy()
    vis, val = out["visits"].cpu().numpy(), out["value"].cpu().numpy()
    lengths = g["lengths"]
    for t in range(T):
        live = t < lengths                                         # env b is recorded for its first lengths[b] moves
        assert np.array_equal(recd[t], live)
        got = actor.keep_rep_inputs[t].cpu().numpy()
        # bit-exact rep-net input for every game still running (a finished game's trajectory stops growing in the reference, :205)
        assert np.array_equal(got[live], g["rep_inputs"][t][live]), f"move {t}: rep-net input differs from _prepare_mcts_input"
    for b in range(B):
        n = int(lengths[b])
        assert np.array_equal(act[:n, b], g[f"t{b}_actions"][32:]), "actions = CDF pick of the injected uniforms on the reference's probabilities"
        assert np.array_equal(frames[:n, b], g[f"t{b}_states"][31:])
        assert np.array_equal(rew[:n, b], g[f"t{b}_rewards"][32:])
        assert np.array_equal(vis[:n, b].astype(np.float32), g[f"t{b}_visits"][32:])
        assert np.array_equal(val[:n, b], g[f"t{b}_values"][32:])
    assert np.array_equal(out["initial_gray"].cpu().numpy()[:, 0], np.stack([g[f"t{b}_states"][0, 0] for b in range(B)]))


def test_sample_actions_matches_reference_probabilities(golden_dir):
    import os
    from muzero_breakout_b200 import _lib
    from oracle import acting_oracle as A
    g = np.load(os.path.join(golden_dir, "acting.npz"))
    visits = torch.from_numpy(g["s_visits"]).cuda()
    B, seed, step = visits.shape[0], 1234, 9
    for k, temp in enumerate(g["s_temps"]):
        action = torch.empty(B, dtype=torch.int64, device="cuda")
        probs = torch.empty((B, 3), dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib().mz_sample_actions(B, visits.data_ptr(), float(temp), seed, step, action.data_ptr(), None, probs.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
        want = g[f"s_probs{k}"]                                    # visit_counts ** (1/T) / sum, evaluated by the reference's torch expression
        assert np.allclose(probs.cpu().numpy(), want, atol=2e-6, rtol=0)
        u = np.array([(rng_u32(seed, b, step) >> 8) / 16777216.0 for b in range(B)], np.float32)
        pick = np.array([A.pick(u[b], want[b]) for b in range(B)])
        near = (np.abs(u - want[:, 0]) < 1e-5) | (np.abs(u - want[:, 0] - want[:, 1]) < 1e-5)       # powf last-bit differences at a CDF boundary
        assert np.array_equal(action.cpu().numpy()[~near], pick[~near]) and near.mean() < 0.01
