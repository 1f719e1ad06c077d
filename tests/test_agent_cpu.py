"""Host logic of the learner-side drop-in on the CPU (no kernel runs): muzero-breakout_b200/src/agent.py mirrors the reference's MuZeroAgent
(src/networks.py:245-350) -- same parameters from the same seed, same outputs on the torch path every bridge falls back to when it does not
qualify (CPU tensors, eval mode, no_grad) -- and train.accelerate_agent leaves a reference-shaped module's results unchanged there."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from muzero_breakout_b200.src.networks import DEFAULT_MODEL_CFG
from oracle.networks import OracleAgent


def _rollout(agent, frames, actions):
    h = agent.create_hidden_state_root(frames)
    p, v = agent.evaluate_state(h)
    planes = torch.nn.functional.one_hot(actions, 3).float().view(-1, 3, 1, 1).expand(-1, -1, 4, 5)
    h2, r = agent.hidden_state_transition(h, planes)
    return h, p, v, h2, r


def test_dropin_agent_equals_the_reference_shaped_agent_on_the_torch_path():
    from muzero_breakout_b200 import train, train_layers
    from muzero_breakout_b200.src.agent import MuZeroAgent
    cfg = dict(DEFAULT_MODEL_CFG, learning_rate=2e-4, device="cpu")
    torch.manual_seed(3)
    ref = OracleAgent(cfg, device="cpu")
    torch.manual_seed(3)
    ours = MuZeroAgent(cfg)
    assert ours.device == "cpu" or not torch.cuda.is_available()
    ours = ours.cpu()
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    assert all(torch.equal(a, b) for a, b in zip(ours.state_dict().values(), ref.state_dict().values()))
    g = torch.Generator().manual_seed(5)
    frames = torch.rand(3, 64, 16, 20, generator=g)
    actions = torch.randint(0, 3, (3,), generator=g)
    for mode in ("train", "eval"):
        getattr(ours, mode)(); getattr(ref, mode)()
        a, b = _rollout(ours, frames, actions), _rollout(ref, frames, actions)
        for u, v in zip(a, b):
            assert u.shape == v.shape and torch.allclose(u, v, rtol=1e-5, atol=1e-6), mode
    # no bridge takes a CPU tensor, whatever the mode
    x = torch.rand(2, 256, 4, 5)
    ours.train()
    assert not train.trunk_supported(ours.pred_net.res_blocks, x)
    assert not train_layers.convblock_supported(ours.dyn_net.conv_block, x, torch.zeros(2, 3, 4, 5))
    assert not train_layers.scale_supported(x) and not train_layers.pool_supported(ours.rep_net.avg_pool, x)
    assert not train_layers.flatten_linear_supported(ours.pred_net.policy_head[2], torch.rand(2, 128, 4, 5))
    # the dynamics network accepts the reference's concatenated input and the (latent, planes) pair
    ours.eval()
    planes = torch.nn.functional.one_hot(actions[:2], 3).float().view(-1, 3, 1, 1).expand(-1, -1, 4, 5)
    with torch.no_grad():
        h1, r1 = ours.dyn_net(torch.cat([x, planes], dim=1))
        h2, r2 = ours.dyn_net(x, planes)
    assert torch.equal(h1, h2) and torch.equal(r1, r2)


def test_accelerate_agent_keeps_a_cpu_module_on_its_own_ops():
    from muzero_breakout_b200 import train
    cfg = dict(DEFAULT_MODEL_CFG)
    torch.manual_seed(4)
    agent = OracleAgent(cfg, device="cpu")
    plain = copy.deepcopy(agent)
    agent.optimizer = torch.optim.Adam(agent.parameters(), lr=2e-4, weight_decay=1e-4)
    train.accelerate_agent(agent)
    assert isinstance(agent.optimizer, torch.optim.Adam)                # the flat-buffer Adam needs CUDA parameters
    g = torch.Generator().manual_seed(6)
    frames = torch.rand(2, 64, 16, 20, generator=g)
    actions = torch.randint(0, 3, (2,), generator=g)
    agent.train(); plain.train()
    for u, v in zip(_rollout(agent, frames, actions), _rollout(plain, frames, actions)):
        assert torch.allclose(u, v, rtol=1e-5, atol=1e-6)
