"""End-to-end MCTSSearchVec.search on the GPU (networks + tree kernels + CUDA graph) against the oracle.

Parity statement (BASELINE.json north_star): root visit counts identical GIVEN IDENTICAL NETWORK OUTPUTS.
Two independently computed forward passes differ in the last bits, which is enough to flip near-tied
pUCT choices, so the search is stepped one simulation at a time: the network
outputs our kernels produced (reward, value, priors: fp32 tensors) are fed to the CPU tree oracle, and
every selection, the visit counts and the root value must be identical; the network outputs themselves
are checked against the fp32 torch oracle evaluated on the same (parent latent, action) to 1e-5."""
import numpy as np
import pytest
import torch

import oracle
from common import dirichlet_noise, perturb_bn
from oracle.networks import OracleAgent

pytestmark = pytest.mark.gpu

CFG = {"num_simulations": 50, "actions": [0, 1, 2], "latent_resolution": [4, 5],
       "search": {"c1": 1.25, "c2": 19652.0, "discount_factor": 0.985, "mcts_name": "MCTSSearchVec"}}


def make(agent, precision, **kw):
    from muzero_breakout_b200.src.mcts import MCTSSearchVec
    cfg = dict(CFG, model=agent.cfg)
    cfg["search"] = dict(CFG["search"], precision=precision, **kw)
    return MCTSSearchVec(cfg, agent, None)


@pytest.fixture(scope="module")
def agent():
    torch.manual_seed(0)
    a = OracleAgent()
    perturb_bn(a, 1)
    a.eval_mode()
    return a


def _stepped(m, hidden, noise, seed):
    """Drive the plan manually (no graph), recording what the networks produced at every simulation."""
    from muzero_breakout_b200.src.mcts import _SearchPlan
    nets = m.packed_networks()
    B = hidden.shape[0]
    plan = _SearchPlan(nets, B, m.num_simulations, m.c1, m.c2, m.discount, m.noise_weight, use_graph=False)
    t = plan.tree
    plan.hidden.copy_(hidden); plan.noise.copy_(noise)
    plan.root_prog.run()
    rec = [dict(v=plan.value.cpu().numpy().copy(), pi=plan.pi.cpu().numpy().copy())]
    t.root(plan.value, plan.pi, plan.noise, m.noise_weight, seed)
    for sim in range(m.num_simulations):
        sel = dict(parent=t.leaf_parent.cpu().numpy().copy(), action=t.leaf_action.cpu().numpy().copy(), leaf=t.leaf_slot.cpu().numpy().copy())
        plan.sim_prog.run()
        sel.update(r=plan.reward.cpu().numpy().copy(), v=plan.value.cpu().numpy().copy(), pi=plan.pi.cpu().numpy().copy())
        rec.append(sel)
        t.step(sim, plan.reward, plan.value, plan.pi, seed)
    return t.out_value.cpu().numpy().copy(), t.out_visits.cpu().numpy().copy(), rec, plan


@pytest.mark.parametrize("precision,B", [("f32", 5), ("bf16", 24), ("bf16", 333), ("f16", 40)])
def test_search_lockstep_identical_visits(agent, precision, B):
    m = make(agent, precision)
    g = torch.Generator().manual_seed(B)
    with torch.no_grad():
        hidden = agent.create_hidden_state_root(torch.rand(min(B, 8), 64, 16, 20, generator=g))
    hidden = hidden.repeat((B + 7) // 8, 1, 1, 1)[:B] * (0.9 + 0.1 * torch.rand(B, 1, 1, 1, generator=g))
    noise, seed = dirichlet_noise(B, 7), 1234
    value, visits, rec, plan = _stepped(m, hidden, noise, seed)
    tree = oracle.TreeOracle(B, 50, 1.25, 19652.0, 0.985, seed)
    parent, action, leaf = tree.root(rec[0]["v"], rec[0]["pi"], noise, m.noise_weight)
    for s in range(50):
        if s > 0:
            parent, action, leaf = tree.select()
        r = rec[s + 1]
        assert np.array_equal(parent, r["parent"]) and np.array_equal(action, r["action"]) and np.array_equal(leaf, r["leaf"]), f"selection differs at sim {s}"
        tree.backup(r["r"], r["v"], r["pi"])
    ovalue, ovisits = tree.results()
    assert np.array_equal(visits, ovisits), "visit counts differ"
    assert np.array_equal(value.view(np.uint32), ovalue.view(np.uint32)), "root value not bit-identical"
    assert np.all(visits.sum(1) == 50)

    # the same search through the public call (CUDA graph) returns the same answer, twice
    for _ in range(2):
        v2, n2 = m.search(hidden, torch.ones(B, 3), 0, noise=noise, seed=seed)
        assert v2.device.type == "cpu" and n2.dtype == torch.int64 and v2.dtype == torch.float32
        assert np.array_equal(n2.numpy(), visits) and np.array_equal(v2.numpy().view(np.uint32), value.view(np.uint32))
    # a different stream key changes tie-breaks (the graph reads the key from device memory)
    v3, n3 = m.search(hidden, torch.ones(B, 3), 0, noise=noise, seed=seed + 1)
    assert n3.sum().item() == 50 * B

    if precision == "f32":
        # network outputs along the searched paths vs the fp32 torch oracle on the same inputs
        lat = [{0: hidden[b]} for b in range(B)]
        with torch.no_grad():
            pol, val = agent.evaluate_state(hidden)
        assert np.allclose(rec[0]["pi"], torch.softmax(pol, 1).numpy(), rtol=0, atol=1e-5)
        assert np.allclose(rec[0]["v"], agent.inverted_softmax_expectation(val).numpy(), rtol=0, atol=1e-5 * 25)
        for s in range(50):
            r = rec[s + 1]
            h = torch.stack([lat[b][int(r["parent"][b])] for b in range(B)])
            planes = torch.zeros(B, 3, 4, 5); planes[torch.arange(B), torch.as_tensor(r["action"]).long()] = 1
            with torch.no_grad():
                h2, rew = agent.hidden_state_transition(h, planes)
                pol, val = agent.evaluate_state(h2)
            for b in range(B):
                lat[b][int(r["leaf"][b])] = h2[b]
            assert np.allclose(r["pi"], torch.softmax(pol, 1).numpy(), rtol=0, atol=2e-5), f"sim {s} priors"
            assert np.allclose(r["r"], agent.inverted_softmax_expectation(rew).numpy(), rtol=0, atol=5e-4), f"sim {s} reward"
            assert np.allclose(r["v"], agent.inverted_softmax_expectation(val).numpy(), rtol=0, atol=5e-4), f"sim {s} value"


def test_mutable_attributes_and_repacking(agent):
    """noise_weight (train_torch.py:135) and mu_zero (:449,451) are reassigned by the caller; a
    load_state_dict into the same module (:361-367) must trigger a re-pack."""
    import copy
    m = make(agent, "bf16")
    B = 6
    hidden = torch.rand(B, 256, 4, 5, generator=torch.Generator().manual_seed(1))
    noise = dirichlet_noise(B, 3)
    v1, n1 = m.search(hidden, None, 0, noise=noise, seed=5)
    m.noise_weight = 0.1
    v2, n2 = m.search(hidden, None, 0, noise=noise, seed=5)
    assert n2.sum() == 50 * B
    other = copy.deepcopy(agent)
    with torch.no_grad():
        other.pred_net.value_head[2].bias[:3].add_(4.0)      # favour the negative supports (a uniform shift would cancel in the softmax)
    m.noise_weight = 0.175
    m.mu_zero = other
    v3, n3 = m.search(hidden, None, 0, noise=noise, seed=5)
    assert not np.allclose(v3.numpy(), v1.numpy())
    other.load_state_dict(agent.state_dict())
    v4, n4 = m.search(hidden, None, 0, noise=noise, seed=5)
    assert np.array_equal(v4.numpy(), v1.numpy()) and np.array_equal(n4.numpy(), n1.numpy())
    # an in-place update of one parameter of the same module (an optimizer step): seen through its version counter on the next call
    with torch.no_grad():
        other.pred_net.value_head[2].bias[:3].add_(4.0)
    v5, n5 = m.search(hidden, None, 0, noise=noise, seed=5)
    assert np.array_equal(v5.numpy(), v3.numpy()) and not np.allclose(v5.numpy(), v1.numpy())


@pytest.mark.parametrize("B", [24, 70])
def test_repeated_graph_replays_are_deterministic(agent, B):
    """The latency-mode networks keep a launch epoch in their hand-off buffers instead of resetting them (csrc/conv_lat.cu):
    many replays of the same captured search with the same noise and seed must return bit-identical results, and a replay
    after a different input must not see anything stale."""
    m = make(agent, "bf16", output_device="cuda")
    g = torch.Generator().manual_seed(B)
    h1, h2 = torch.rand(B, 256, 4, 5, generator=g), torch.rand(B, 256, 4, 5, generator=g)
    noise = dirichlet_noise(B, seed=3)
    v1, n1 = m.search(h1, None, 0, noise=noise, seed=11)
    v2, n2 = m.search(h2, None, 0, noise=noise, seed=11)
    assert int(n1.sum()) == B * CFG["num_simulations"] and not torch.equal(v1, v2)
    for rep in range(12):
        va, na = m.search(h1, None, 0, noise=noise, seed=11)
        assert torch.equal(va, v1) and torch.equal(na, n1), f"replay {rep} of input 1 differs"
        vb, nb = m.search(h2, None, 0, noise=noise, seed=11)
        assert torch.equal(vb, v2) and torch.equal(nb, n2), f"replay {rep} of input 2 differs"


def test_full_size_search_properties(agent):
    """BASELINE.json configs[2]'s size (4096 roots x 50 simulations, the default f16 mode) through size-independent properties:
    visit counts sum to num_simulations in every tree, finite values, bit-identical repeat calls, and batch independence -- a tree's
    result depends only on its own latent, noise and tie-break stream, so the first 2560 trees of the 4096-root search equal a 2560-root
    search on the same inputs (both sizes run the full-width tcgen05 trunk, whose per-sample arithmetic does not depend on the batch)."""
    B, S = 4096, 50
    m = make(agent, "f16", output_device="cuda")
    g = torch.Generator().manual_seed(21)
    hidden = torch.rand(B, 256, 4, 5, generator=g)
    noise = dirichlet_noise(B, 9)
    v1, n1 = m.search(hidden, None, 0, noise=noise, seed=77)
    assert n1.dtype == torch.int64 and tuple(n1.shape) == (B, 3) and bool((n1.sum(1) == S).all())
    assert bool(torch.isfinite(v1).all()) and bool((n1 >= 0).all())
    assert int((n1.max(1).values == S).sum()) < B // 2, "degenerate trees: every simulation went down one action"
    v2, n2 = m.search(hidden, None, 0, noise=noise, seed=77)          # CUDA-graph replay: bit-identical
    assert torch.equal(v1, v2) and torch.equal(n1, n2)
    v3, n3 = m.search(hidden, None, 0, noise=noise, seed=78)          # another tie-break stream: different trees, same invariants
    assert bool((n3.sum(1) == S).all()) and not torch.equal(n1, n3)
    sub = 2560
    vs, ns = m.search(hidden[:sub].contiguous(), None, 0, noise=noise[:sub], seed=77)
    assert torch.equal(ns, n1[:sub]) and torch.equal(vs, v1[:sub]), "a tree's search must not depend on the other trees of the batch"
