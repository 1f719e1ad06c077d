"""GPU parity tests of the tree kernels (csrc/tree.cu through mz_tree_root / mz_tree_step) against the
golden records of the UNMODIFIED reference MCTSSearchVec.search (tests/golden/mcts_*.npz) and against
the CPU tree oracle, in lock step: the recorded / generated network outputs are fed to both sides, so
"identical network outputs" means the same fp32 tensors.  Visit counts identical, root value bit-exact,
every selected (parent, action, leaf slot) identical at every simulation."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "mcts_fake.npz"))
    for i in range(len(g["modes"])):
        yield f"fake{i}_{g['modes'][i]}", {k[len(f"c{i}_"):]: g[k] for k in g.files if k.startswith(f"c{i}_")}
    r = np.load(os.path.join(golden_dir, "mcts_real.npz"))
    rec = {k: r[k] for k in r.files}
    rec["meta"] = np.array([rec["meta"][0], rec["meta"][1], rec["meta"][2], 0.175, 1.25, 19652.0])
    yield "real", rec


def _cuda(x, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(x)).to("cuda", dtype)


def test_tree_lockstep_vs_reference_golden(golden_dir):
    from muzero_breakout_b200.src.mcts import TreeBuffers
    for name, rec in _cases(golden_dir):
        B, S, seed, w, c1, c2 = rec["meta"]
        B, S, seed = int(B), int(S), int(seed)
        t = TreeBuffers(B, S, c1, c2, 0.985, "cuda", latent_bytes=16)
        # fake latent store: 4 int32 words (tree, slot, tree^slot, 7) so the gather can be checked
        store = t.latent_store.view(torch.int32).view(B, t.nodes, 4)
        bb, ss = torch.meshgrid(torch.arange(B), torch.arange(t.nodes), indexing="ij")
        store.copy_(torch.stack([bb, ss, bb ^ ss, torch.full_like(bb, 7)], dim=-1).to(torch.int32))
        t.root(_cuda(rec["v_root"]), _cuda(rec["pi_root"]), _cuda(rec["noise"]), w, seed)
        for s in range(S):
            assert np.array_equal(t.leaf_parent.cpu().numpy(), rec["parent"][s]), f"{name}: parent differs at sim {s}"
            assert np.array_equal(t.leaf_action.cpu().numpy(), rec["action"][s]), f"{name}: action differs at sim {s}"
            assert np.array_equal(t.leaf_slot.cpu().numpy(), rec["leaf"][s]), f"{name}: leaf slot differs at sim {s}"
            got = t.dyn_in.view(torch.int32).view(B, 4).cpu().numpy()
            assert np.array_equal(got[:, 0], np.arange(B)) and np.array_equal(got[:, 1], rec["parent"][s]), f"{name}: gather wrong at sim {s}"
            t.step(s, _cuda(rec["reward"][s]), _cuda(rec["leaf_value"][s]), _cuda(rec["pi"][s]), seed)
        visits, value = t.out_visits.cpu().numpy(), t.out_value.cpu().numpy()
        assert np.array_equal(visits, rec["visits"]), f"{name}: visit counts differ"
        assert np.array_equal(value.view(np.uint32), rec["value"].view(np.uint32)), f"{name}: root value not bit-identical"


@pytest.mark.parametrize("B,S,mode", [(3000, 50, "mild"), (777, 100, "optimistic"), (4096, 50, "ties"), (1, 1, "mild"), (2, 2, "ties"),
                                      (33, 300, "optimistic")])
def test_tree_large_batch_vs_oracle(B, S, mode):
    """Thousands of trees, random network outputs, S=100 exercises node slots beyond the shared-memory
    stage; 'ties' feeds constant outputs so that every pUCT call is an exact tie."""
    from muzero_breakout_b200.src.mcts import TreeBuffers
    g = torch.Generator().manual_seed(B + S)
    seed = 99 + S
    c1, c2, disc, w = 1.25, 19652.0, 0.985, 0.175

    def outputs():
        if mode == "ties":
            return torch.zeros(B), torch.full((B,), 0.25), torch.full((B, 3), 1 / 3)
        r = (torch.rand(B, generator=g) - 0.4) * (0.2 if mode == "mild" else 1.0)
        v = (torch.rand(B, generator=g) - 0.5) * (0.4 if mode == "mild" else 3.0) + (0.0 if mode == "mild" else 1.0)
        return r, v, torch.softmax(torch.randn(B, 3, generator=g), dim=1)

    _, v0, pi0 = outputs()
    noise = torch.distributions.Dirichlet(torch.full((3,), 0.25)).sample((B,)) if mode != "ties" else torch.full((B, 3), 1 / 3)
    t = TreeBuffers(B, S, c1, c2, disc, "cuda")
    t.depth_hist = torch.zeros(S + 1, dtype=torch.int32, device="cuda")
    orc = oracle.TreeOracle(B, S, c1, c2, disc, seed)
    t.root(v0.cuda(), pi0.cuda(), noise.cuda(), w, seed)
    parent, action, leaf = orc.root(v0, pi0, noise, w)
    for s in range(S):
        if s > 0:
            parent, action, leaf = orc.select()
        assert np.array_equal(t.leaf_parent.cpu().numpy(), parent), f"parent differs at sim {s}"
        assert np.array_equal(t.leaf_action.cpu().numpy(), action), f"action differs at sim {s}"
        assert np.array_equal(t.leaf_slot.cpu().numpy(), leaf), f"leaf differs at sim {s}"
        r, v, pi = outputs()
        t.step(s, r.cuda(), v.cuda(), pi.cuda(), seed)
        orc.backup(r, v, pi)
    ovalue, ovisits = orc.results()
    assert np.array_equal(t.out_visits.cpu().numpy(), ovisits)
    assert np.array_equal(t.out_value.cpu().numpy().view(np.uint32), ovalue.view(np.uint32))
    assert np.all(ovisits.sum(1) == S)
    hist = t.depth_hist.cpu().numpy()
    assert hist.sum() == B * S and (S < 3 or hist[1:].sum() > 0)
