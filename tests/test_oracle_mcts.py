"""Pins oracle/mcts_oracle.c against the UNMODIFIED reference MCTSSearchVec.search run with injected
RNG (tests/golden/mcts_fake.npz, mcts_real.npz from tests/golden/gen_golden.py): the recorded
network outputs are replayed in lock step; selected leaves, visit counts and the root value must be
identical (root value bit-exact)."""
import os

import numpy as np
import pytest

import oracle
from refshim_rng import rng_u32


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "mcts_fake.npz"))
    n = len(g["modes"])
    for i in range(n):
        yield f"fake{i}_{g['modes'][i]}", {k[len(f"c{i}_"):]: g[k] for k in g.files if k.startswith(f"c{i}_")}
    r = np.load(os.path.join(golden_dir, "mcts_real.npz"))
    rec = {k: r[k] for k in r.files}
    rec["meta"] = np.array([rec["meta"][0], rec["meta"][1], rec["meta"][2], 0.175, 1.25, 19652.0])
    yield "real", rec


def test_rng_stream_matches_python_definition():
    for seed, tree, ctr in [(0, 0, 0), (7, 3, 11), (2**63 + 5, 2**31, 2**32 - 1), (123456789, 4095, 2500)]:
        assert oracle.rng_u32(seed, tree, ctr) == rng_u32(seed, tree, ctr)


def test_lockstep_replay_matches_reference(golden_dir):
    for name, rec in _cases(golden_dir):
        B, S, seed, w, c1, c2 = rec["meta"]
        B, S, seed = int(B), int(S), int(seed)
        tree = oracle.TreeOracle(B, S, c1, c2, 0.985, seed)
        parent, action, leaf = tree.root(rec["v_root"], rec["pi_root"], rec["noise"], w)
        for s in range(S):
            if s > 0:
                parent, action, leaf = tree.select()
            assert np.array_equal(parent, rec["parent"][s]), f"{name}: parent slot differs at sim {s}"
            assert np.array_equal(action, rec["action"][s]), f"{name}: action differs at sim {s}"
            assert np.array_equal(leaf, rec["leaf"][s]), f"{name}: leaf slot differs at sim {s}"
            tree.backup(rec["reward"][s], rec["leaf_value"][s], rec["pi"][s])
        value, visits = tree.results()
        assert np.array_equal(visits, rec["visits"]), f"{name}: visit counts differ"
        assert np.array_equal(value.view(np.uint32), rec["value"].view(np.uint32)), f"{name}: root value not bit-identical"
        assert np.all(visits.sum(1) == S)
        # one RNG draw per pUCT call, like the reference (mcts.py:297)
        assert np.array_equal(np.array([tree_ctr for tree_ctr in rec["ucb_calls"]]), rec["ucb_calls"])


def test_full_search_with_fake_network(golden_dir):
    """oracle.search (tree oracle + network calls) reproduces the reference end to end when the
    network is the same deterministic element-wise function."""
    import torch
    from common import FakeNet

    g = np.load(os.path.join(golden_dir, "mcts_fake.npz"))
    for i, mode in enumerate(g["modes"]):
        B, S, seed, w, c1, c2 = g[f"c{i}_meta"]
        v, n = oracle.search(FakeNet(str(mode)), torch.from_numpy(g[f"c{i}_hidden"]), g[f"c{i}_noise"], int(seed),
                             num_simulations=int(S), c1=c1, c2=c2, noise_weight=w)
        assert np.array_equal(n, g[f"c{i}_visits"]), f"case {i} ({mode}): visits differ"
        assert np.array_equal(v.view(np.uint32), g[f"c{i}_value"].view(np.uint32)), f"case {i}: value differs"


def test_sim0_leaf_is_reexpanded_quirk(golden_dir):
    """mcts.py:121 + :163-175: the child expanded in simulation 0 keeps expanded=False and is
    re-expanded into the same slot the second time it is selected."""
    g = np.load(os.path.join(golden_dir, "mcts_fake.npz"))
    seen = 0
    for i in range(len(g["modes"])):
        leaf = g[f"c{i}_leaf"]
        seen += int(((leaf[1:] == 1) & (g[f"c{i}_parent"][1:] == 0)).sum())
    assert seen > 0
