"""oracle/replay_oracle.py against the golden vectors of the reference's ReplayBuffer (tests/golden/replay.npz)."""
import os

import numpy as np
import pytest

from oracle.replay_oracle import ReplayOracle

FIELDS = ("past_actions", "future_actions", "states", "rewards", "visit_counts", "values", "value_buffer")


def load_replay_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "replay.npz"))
    cap, K, hist, n_sum = (int(x) for x in g["meta"])
    trajs = [{k: g[f"t{i}_{k}"] for k in ("init", "frames", "action", "reward", "visits", "value")} for i in range(len(g["lengths"]))]
    return g, dict(cap=cap, K=K, hist=hist, n_sum=n_sum, discount=float(g["discount"])), trajs


def test_replay_oracle_matches_reference_goldens(golden_dir):
    g, p, trajs = load_replay_golden(golden_dir)
    rb = ReplayOracle(p["hist"], p["K"], p["cap"], p["discount"], p["n_sum"])
    snap_after = list(g["snap_after"])
    for ti, tr in enumerate(trajs):
        rb.save(tr["init"], tr["frames"], tr["action"], tr["reward"], tr["visits"], tr["value"])
        assert len(rb) == int(g["length_after"][ti])
        if ti in snap_after:
            tag = f"s{snap_after.index(ti)}_"
            idx = np.arange(len(rb))
            for f in FIELDS:
                want = g[tag + f]
                got = rb.batch(f, idx)
                assert got.shape == want.shape, f
                assert np.array_equal(got.astype(want.dtype), want), f            # bit-exact, fp32 value targets included
            assert np.array_equal(rb.batch("states", g[tag + "perm"]), g[tag + "states"][g[tag + "perm"]])
            assert rb.reward_sums() == list(g[tag + "reward_sums"])
            assert [x["reward_sum"] for x in rb.samples] == list(g[tag + "reward_sums_all"])


def test_replay_golden_covers_both_value_target_branches(golden_dir):
    g, p, trajs = load_replay_golden(golden_dir)
    # trajectories shorter than K give no sample, K..K+9 only the tail-sum branch, longer ones both
    assert min(g["lengths"]) < p["K"] and max(g["lengths"]) == 261
    assert int(g["length_after"][0]) == 0 and int(g["length_after"][1]) == 1
    assert int(g["length_after"][-1]) == p["cap"]
