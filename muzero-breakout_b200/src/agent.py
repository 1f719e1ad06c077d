"""Drop-in for the reference's `src/networks.py` MuZeroAgent (:245-350) -- the LEARNER side.

Same constructor (`MuZeroAgent(cfg["model"])`, train_torch.py:85-87 through get_class("src.networks", agent_name)), same module tree and
state_dict keys (parameters are created in the reference's order, so the same torch.manual_seed gives the same
initial weights), same methods (`create_hidden_state_root`, `hidden_state_transition`, `evaluate_state`, `_scale_state`, `train_mode`,
`eval_mode`, `.optimizer`, `.device`), so the reference's `_training_stage` / `_k_step_rollout` (train_torch.py:369-528) run on it unchanged.

What runs where in a training step (`loss.backward()` included):
  * the ResidualBlock runs of the dynamics and prediction networks (2 x 14 blocks, K = 5 unroll steps: 280 of the ~330 convolutions of a
    step) -- forward (fp16 operands), data gradient and weight gradient (bf16 gradients) on tcgen05 and the training-mode BatchNorm kernels, through
    train.trunk_forward;
  * the optimizer -- train.Adam (one mz_adam launch over flat buffers);
  * the representation network's 256-channel ResidualBlocks (3 at 16x20, 3 at 8x10) on the same kernels (MZB_TRAIN_ANY_HW=0: torch ops);
  * the representation network's stem convolutions, 128-channel blocks and pools, the dynamics ConvBlock with its action planes, the three
    head ConvBlocks + Linear heads and `_scale_state` -- library kernels as well (train_layers.py: tcgen05 convolution / data / weight
    gradients, csrc/train_layers.cu for the pools, Linear heads, `_scale_state` and the action-plane channels; MZB_TRAIN_LAYERS=0: torch
    ops).  What is left to torch inside a step is glue: torch.cat / stack / one_hot of the rollout, autograd's gradient accumulation adds,
    the 16-bit weight re-packs.
In eval mode / under no_grad the module runs plain torch ops; acting does not call it at all (MCTSSearchVec packs its state_dict).
There is no CPU path for the accelerated parts: on a CPU tensor the module is an ordinary torch module.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import train as _train
from .. import train_layers as _layers

_ACTS = {"relu": nn.ReLU, "leaky_relu": nn.LeakyReLU, "silu": nn.SiLU, "gelu": nn.GELU}      # utils.py:99-108


class ConvBlock(nn.Module):
    def __init__(self, activation: str, in_ch: int, out_ch: int, stride: int = 1, kernel_size: int = 3, padding: int = 1):
        super().__init__()
        self.conv = nn.Conv2d(in_ch, out_ch, kernel_size, stride, padding)
        self.bn = nn.BatchNorm2d(out_ch)
        self.act = _ACTS[activation]()

    def forward(self, x):
        if _layers.convblock_supported(self, x):
            return _layers.convblock_forward(self, x)
        return self.act(self.bn(self.conv(x)))


class ResidualBlock(nn.Module):
    def __init__(self, in_ch: int, activation: str):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, in_ch, 3, 1, 1)
        self.bn1 = nn.BatchNorm2d(in_ch)
        self.conv2 = nn.Conv2d(in_ch, in_ch, 3, 1, 1)
        self.bn2 = nn.BatchNorm2d(in_ch)
        self.act = _ACTS[activation]()

    def forward(self, x):
        if _train.trunk_supported([self], x):
            return _train.trunk_forward([self], x)
        y = self.act(self.bn1(self.conv1(x)))
        return self.act(self.bn2(self.conv2(y)) + x)


def _is_resblock(m) -> bool:
    """this module's ResidualBlock or the reference's (networks.py:19-35): accelerate_agent() binds these forwards to reference modules"""
    return all(hasattr(m, a) for a in ("conv1", "bn1", "conv2", "bn2"))


def _run_blocks(blocks, x):
    """a run of ResidualBlocks: one library call when the kernels take it, the modules one by one otherwise"""
    if _train.trunk_supported(blocks, x):
        return _train.trunk_forward(blocks, x)
    for b in blocks:
        x = b(x)
    return x


class RepresentationNetwork(nn.Module):
    def __init__(self, cfg: dict, in_ch: int):
        super().__init__()
        c0, c1 = cfg["latent_channels"]
        act = cfg["representation_network"]["activation"]
        n0, n1, n2 = cfg["representation_network"]["num_res_blocks"]
        self.avg_pool = nn.AvgPool2d(kernel_size=(2, 2), stride=2)
        self.blocks = nn.ModuleList([])
        self.blocks.append(nn.Conv2d(in_ch, c0, 3, 1, 1))
        for _ in range(n0):
            self.blocks.append(ResidualBlock(c0, act))
        self.blocks.append(nn.Conv2d(c0, c1, 3, 1, 1))
        for _ in range(n1):
            self.blocks.append(ResidualBlock(c1, act))
        self.blocks.append(self.avg_pool)
        for _ in range(n2):
            self.blocks.append(ResidualBlock(c1, act))
        self.blocks.append(self.avg_pool)

    def forward(self, state):
        if state.is_cuda and not _layers.conv_supported(self.blocks[0], state):
            # (torch-op path) channels_last activations from here on: cuDNN then runs its NHWC kernels without the NCHW <-> NHWC conversion pair around every
            # convolution (5.8 ms of a 512-sample training step), and the library trunks take / return their channels-last layout without a copy
            state = state.contiguous(memory_format=torch.channels_last)
        run = []                                       # consecutive ResidualBlocks go through the library as one call (no layout / precision
        for m in self.blocks:                          # round trip between them); the stems and pools through their own bridges
            if _is_resblock(m):
                run.append(m)
                continue
            if run:
                state, run = _run_blocks(run, state), []
            if isinstance(m, nn.Conv2d) and _layers.conv_supported(m, state):
                state = _layers.conv_forward(m, state)
            elif _layers.pool_supported(m, state):
                state = _layers.pool_forward(state)
            else:
                state = m(state)
        if run:
            state = _run_blocks(run, state)
        return state


class DynamicsNetwork(nn.Module):
    def __init__(self, cfg: dict, in_ch: int, latent_resolution):
        super().__init__()
        self.num_res_blocks = cfg["dynamics_network"]["num_res_blocks"]
        act = cfg["dynamics_network"]["activation"]
        self.conv_block = ConvBlock(act, in_ch + cfg["dynamics_network"]["num_actions"], in_ch, 1)
        self.res_blocks = nn.ModuleList([ResidualBlock(in_ch, act) for _ in range(self.num_res_blocks)])
        self.state_head = nn.Sequential()
        self.reward_head = nn.Sequential(ConvBlock(act, in_ch, in_ch, 1, kernel_size=1, padding=0), nn.Flatten(1, -1),
                                         nn.Linear(in_ch * latent_resolution[0] * latent_resolution[1], cfg["num_supports"]))

    def forward(self, hidden_state, action_planes=None):
        """hidden_state: the reference's torch.cat([latent, action planes], 1) (networks.py:295) -- or the latent alone with the planes as a
        second argument (MuZeroAgent.hidden_state_transition below: no concatenated copy)"""
        if action_planes is None and hidden_state.shape[1] == self.conv_block.conv.weight.shape[1]:
            n_lat = self.conv_block.bn.num_features
            h, planes = hidden_state[:, :n_lat], hidden_state[:, n_lat:]
        else:
            h, planes = hidden_state, action_planes
        if planes is not None and _layers.convblock_supported(self.conv_block, h, planes):
            x = _layers.convblock_forward(self.conv_block, h, planes)
        else:
            x = self.conv_block(hidden_state if action_planes is None else torch.cat([hidden_state, action_planes], dim=1))
        x = _run_blocks(self.res_blocks, x)
        return x, _layers.head_forward(self.reward_head, x)


class PredictionNetwork(nn.Module):
    def __init__(self, cfg: dict, in_ch: int, latent_resolution):
        super().__init__()
        self.num_res_blocks = cfg["prediction_network"]["num_res_blocks"]
        act = cfg["prediction_network"]["activation"]
        hw = latent_resolution[0] * latent_resolution[1]
        self.res_blocks = nn.ModuleList([ResidualBlock(in_ch, act) for _ in range(self.num_res_blocks)])
        self.policy_head = nn.Sequential(ConvBlock(act, in_ch, in_ch // 2, 1), nn.Flatten(1, -1),
                                         nn.Linear((in_ch // 2) * hw, cfg["prediction_network"]["num_actions"]))
        self.value_head = nn.Sequential(ConvBlock(act, in_ch, in_ch // 2, 1, kernel_size=1, padding=0), nn.Flatten(1, -1),
                                        nn.Linear((in_ch // 2) * hw, cfg["num_supports"]))

    def forward(self, hidden_state):
        x = _run_blocks(self.res_blocks, hidden_state)
        return _layers.head_forward(self.policy_head, x), _layers.head_forward(self.value_head, x)


class MuZeroAgent(nn.Module):
    def __init__(self, cfg: dict):
        super().__init__()
        planes = cfg["state_history_length"] * 2                          # 32 gray frames + 32 action planes (networks.py:248)
        self.device = cfg.get("device", "cuda") if torch.cuda.is_available() else "cpu"
        if self.device == "cuda" or str(self.device).startswith("cuda"):
            self.device = "cuda"                                          # the reference hard-codes it (:249)
        c1 = cfg["latent_channels"][1]
        self.rep_net = RepresentationNetwork(cfg, planes).to(self.device)
        self.dyn_net = DynamicsNetwork(cfg, c1, cfg["latent_resolution"]).to(self.device)
        self.pred_net = PredictionNetwork(cfg, c1, cfg["latent_resolution"]).to(self.device)
        if self.device == "cuda":
            self.optimizer = _train.Adam(self.parameters(), lr=cfg["learning_rate"], weight_decay=0.0001)       # networks.py:268
        else:
            self.optimizer = torch.optim.Adam(self.parameters(), lr=cfg["learning_rate"], weight_decay=0.0001)

    def create_hidden_state_root(self, state: torch.Tensor):
        return self._scale_state(self.rep_net(state.to(self.device)))

    def hidden_state_transition(self, prev_hidden_state: torch.Tensor, action: torch.Tensor):
        if _layers.convblock_supported(self.dyn_net.conv_block, prev_hidden_state, action):
            hidden_state, reward = self.dyn_net(prev_hidden_state, action)        # no concatenated copy: the planes go to their own kernel
        else:
            x = torch.cat([prev_hidden_state, action], dim=1)
            if x.is_cuda:
                x = x.contiguous(memory_format=torch.channels_last)       # see RepresentationNetwork.forward
            hidden_state, reward = self.dyn_net(x)
        return self._scale_state(hidden_state), reward

    def evaluate_state(self, hidden_state: torch.Tensor):
        return self.pred_net(hidden_state)

    def _scale_state(self, hidden_state: torch.Tensor):
        if _layers.scale_supported(hidden_state):
            return _layers.scale_state(hidden_state)
        flat = hidden_state.reshape(hidden_state.shape[0], -1)
        s_min = flat.min(dim=1, keepdim=True)[0].view(-1, 1, 1, 1)
        s_max = flat.max(dim=1, keepdim=True)[0].view(-1, 1, 1, 1)
        return (hidden_state - s_min) / (s_max - s_min + 1e-8)

    def _encode_action(self, action: int):
        return

    def eval_mode(self):
        self.rep_net.eval(); self.dyn_net.eval(); self.pred_net.eval()

    def train_mode(self):
        self.rep_net.train(); self.dyn_net.train(); self.pred_net.train()
