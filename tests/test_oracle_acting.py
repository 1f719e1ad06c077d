"""oracle/acting_oracle.py against tests/golden/acting.npz: one episode of the reference's own RLSystem._run_episode
(train_torch.py:171-233) with injected search outputs and an injected uniform stream for the Categorical draw."""
import os

import numpy as np

import oracle
from oracle import acting_oracle as A
from refshim_rng import rng_u32


def load_acting_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "acting.npz"))
    B, T, seed, env_seed, sims = (int(x) for x in g["meta"])
    return g, B, T, seed, env_seed, sims


def test_acting_oracle_replays_the_reference_episode(golden_dir):
    g, B, T, seed, env_seed, sims = load_acting_golden(golden_dir)
    temperature, ep_seed = float(g["temperature"]), int(g["ep_seed"])
    assert ep_seed == (seed * 0x9E3779B97F4A7C15 + 7) & ((1 << 64) - 1)
    env = oracle.EnvOracle(B)
    state = g["initial_state"].copy()
    env.ball_dx[:] = g["initial_dx"]; env.ball_dy[:] = -1.0
    cur = oracle.gray(state)
    trajs = [A.TrajectoryOracle(cur[b]) for b in range(B)]
    done = np.zeros(B, np.uint8)
    for t in range(T):
        for b in range(B):
            assert np.array_equal(A.rep_input(trajs[b], cur[b]), g["rep_inputs"][t, b]), f"move {t} env {b}: rep-net input"
        p = A.sample_probs(g["visits"][t], temperature)
        assert np.allclose(p, g["probs"][t], atol=1e-6, rtol=0), f"move {t}: probabilities"
        u = np.array([(rng_u32(ep_seed, b, t) >> 8) / 16777216.0 for b in range(B)], np.float32)
        assert np.array_equal(u, g["u"][t])
        act = np.array([A.pick(u[b], g["probs"][t, b]) for b in range(B)])
        prev_done = done.copy()
        state, r, done, _ = env.step(state, act, done)
        cur = oracle.gray(state)
        for b in range(B):
            if not prev_done[b]:
                trajs[b].add(act[b], cur[b], r[b], g["visits"][t, b], g["value"][t, b])
    assert done.all()
    for b in range(B):
        assert trajs[b].length == int(g["lengths"][b])
        assert np.array_equal(np.array(trajs[b].actions), g[f"t{b}_actions"])
        assert np.array_equal(np.stack(trajs[b].states), g[f"t{b}_states"])
        assert np.array_equal(np.array(trajs[b].rewards, np.float32), g[f"t{b}_rewards"])
        assert np.array_equal(np.stack(trajs[b].visits), g[f"t{b}_visits"])
        assert np.array_equal(np.array(trajs[b].values, np.float32), g[f"t{b}_values"])


def test_sampling_probabilities_match_reference_expression(golden_dir):
    g = np.load(os.path.join(golden_dir, "acting.npz"))
    for k, temp in enumerate(g["s_temps"]):
        p = A.sample_probs(g["s_visits"], float(temp))
        assert np.allclose(p, g[f"s_probs{k}"], atol=1e-6, rtol=0)
        assert np.all(p[g["s_visits"] == 0] == 0)
