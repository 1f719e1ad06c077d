"""Pins oracle/networks.py (fp32 torch restatement) against outputs of the reference MuZeroAgent
(tests/golden/mcts_real.npz: torch.manual_seed(0) random-init weights, perturb_bn(seed 1), eval
mode).  Same seed => bit-identical weights, so outputs agree to fp32 rounding (1e-5 relative)."""
import os

import numpy as np
import pytest
import torch

import oracle
from common import perturb_bn
from oracle.networks import OracleAgent, inverted_softmax_expectation

RTOL = 1e-5


def _close(a, b, name, rtol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    assert err <= rtol, f"{name}: max rel-to-range err {err:.3e} > {rtol}"


@pytest.fixture(scope="module")
def agent():
    torch.manual_seed(0)
    a = OracleAgent()
    perturb_bn(a, 1)
    a.eval_mode()
    return a


@pytest.fixture(scope="module")
def rec(golden_dir):
    g = np.load(os.path.join(golden_dir, "mcts_real.npz"))
    return {k: g[k] for k in g.files}


def test_state_dict_layout_matches_appendix_d(agent):
    sd = agent.state_dict()
    assert sd["rep_net.blocks.0.weight"].shape == (128, 64, 3, 3)
    assert sd["rep_net.blocks.3.weight"].shape == (256, 128, 3, 3)
    assert sd["dyn_net.conv_block.conv.weight"].shape == (256, 259, 3, 3)
    assert sd["dyn_net.reward_head.0.conv.weight"].shape == (256, 256, 1, 1)
    assert sd["dyn_net.reward_head.2.weight"].shape == (11, 5120)
    assert sd["pred_net.policy_head.0.conv.weight"].shape == (128, 256, 3, 3)
    assert sd["pred_net.policy_head.2.weight"].shape == (3, 2560)
    assert sd["pred_net.value_head.0.conv.weight"].shape == (128, 256, 1, 1)
    assert sd["pred_net.value_head.2.weight"].shape == (11, 2560)
    n = lambda p: sum(v.numel() for k, v in sd.items() if k.startswith(p) and "running" not in k and "num_batches" not in k)
    assert (n("rep_net."), n("dyn_net."), n("pred_net.")) == (8047488, 17256715, 16900878)


def test_networks_match_reference_outputs(agent, rec):
    with torch.no_grad():
        rep_in = torch.from_numpy(rec["rep_in"])
        _close(agent.rep_net(rep_in), rec["rep_raw"], "rep_net raw")
        hidden = agent.create_hidden_state_root(rep_in)
        _close(hidden, rec["hidden"], "root latent")
        h = torch.from_numpy(rec["hidden"])
        pol, val = agent.evaluate_state(h)
        _close(pol, rec["root_policy_logits"], "root policy logits")
        _close(val, rec["root_value_logits"], "root value logits")
        B = h.shape[0]
        planes = torch.zeros(B, 3, 4, 5)
        planes[torch.arange(B), torch.from_numpy(rec["dyn_actions"])] = 1
        h2, rew = agent.hidden_state_transition(h, planes)
        _close(h2, rec["dyn_h"], "dynamics latent")
        _close(rew, rec["dyn_reward_logits"], "reward logits")
        pol2, val2 = agent.evaluate_state(torch.from_numpy(rec["dyn_h"]))
        _close(pol2, rec["pred_policy_logits"], "policy logits")
        _close(val2, rec["pred_value_logits"], "value logits")
        _close(inverted_softmax_expectation(torch.from_numpy(rec["root_value_logits"])), rec["v_root"], "v_root", 1e-6)


def test_inverse_transform_is_the_reference_formula():
    # utils.py:26-28: sign(x)*((|x| + 0.999)^2 - 1); NOT the true inverse (x=0.0005 -> about -0.001)
    logits = torch.zeros(1, 11); logits[0, 5] = 30.0            # expectation ~ 0 -> |y| ~ 0.002, not 0
    assert abs(abs(float(inverted_softmax_expectation(logits))) - (1 - 0.999 ** 2)) < 1e-5
    x = torch.tensor([[0.0] * 5 + [20.0, 12.4] + [0.0] * 4])
    e = float(torch.sum(torch.softmax(x, -1) * torch.arange(-5, 6).float()))
    assert np.isclose(float(inverted_softmax_expectation(x)), np.sign(e) * ((abs(e) + 0.999) ** 2 - 1), rtol=1e-4)
