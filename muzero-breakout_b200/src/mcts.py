"""Drop-in for the reference's `src/mcts.py` (class MCTSSearchVec :10) on B200.

Same constructor and call signature, so train_torch.py:90-92 can load it by name:

    MCTSSearchVec(cfg, mu_zero, scalar_transforms)
    .search(hidden_state (B,256,4,5), action_mask (B,3), training_iteration)
        -> (value float32 (B,), visit_counts int64 (B,3))           # CPU tensors, like mcts.py:71
    public mutable attributes .noise_weight (train_torch.py:135) and .mu_zero (:449,451)

What runs: the root prediction, then num_simulations x { dynamics + prediction networks on one leaf
per tree (tcgen05 fp16 / bf16, or fp32 CUDA cores with precision="f32") -> fused backup + next selection
(csrc/tree.cu) }, all enqueued on one CUDA stream with no host synchronisation inside the search and,
by default, replayed as one CUDA graph.  Trees are flat per-root arrays in HBM; latents of expanded
nodes live in a preallocated [B][num_simulations+2] store.

Reference quirks reproduced: action_mask and training_iteration are ignored
(mcts.py:124,157), raw Q without min-max statistics (:289), the sim-0 leaf is re-expanded on its second
visit (:121,:163-175), one tie-break draw per pUCT call (:297), root value = sum of 51 terms / 50 (:248).
Randomness: Dirichlet(0.25) root noise (:114) is sampled per call on the device (or passed in with
noise=...); the tie-break draw is the counter-based stream u32(seed, tree, call counter) % count.

Extra, optional: cfg["search"] keys "precision" ("f16" default: within 1e-3 of the fp32 networks / "bf16" / "f32"), "seed", "use_graph",
"output_device" ("cpu" default / "cuda"); search(..., noise=, seed=) for reproducible runs.
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from .networks import F32, BF16, OP_NCHW_IN, PackedNetworks, Program


def _p(t):
    return None if t is None else t.data_ptr()


class TreeBuffers:
    """Device memory of a batch of search trees + thin wrappers over mz_tree_root / mz_tree_step."""

    def __init__(self, B, S, c1, c2, discount, device, latent_bytes=0):
        L = _lib.lib()
        self.B, self.S = int(B), int(S)
        self.device = torch.device(device)
        self.nodes = L.mz_tree_nodes(self.S)
        self.tree_bytes = L.mz_tree_bytes(self.S)
        self.trees = torch.zeros((self.B, self.tree_bytes // 16, 4), dtype=torch.int32, device=self.device)
        s_tab, k_tab = np.zeros(self.S + 1, np.float32), np.zeros(self.S + 1, np.float32)
        _lib.check(L.mz_puct_tables(self.S, float(c1), float(c2), s_tab.ctypes.data, k_tab.ctypes.data))
        self.s_tab = torch.from_numpy(s_tab).to(self.device)
        self.k_tab = torch.from_numpy(k_tab).to(self.device)
        self.discount = float(discount)
        i32 = dict(dtype=torch.int32, device=self.device)
        self.leaf_parent, self.leaf_action, self.leaf_slot = (torch.zeros(self.B, **i32) for _ in range(3))
        self.out_value = torch.zeros(self.B, dtype=torch.float32, device=self.device)
        self.out_visits = torch.zeros((self.B, 3), dtype=torch.int64, device=self.device)
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.depth_hist = None
        self.latent_bytes = int(latent_bytes)
        self.latent_store = self.dyn_in = None
        if latent_bytes:
            self.latent_store = torch.empty((self.B, self.nodes, latent_bytes), dtype=torch.uint8, device=self.device)
            self.dyn_in = torch.empty((self.B, latent_bytes), dtype=torch.uint8, device=self.device)

    def _args(self, sim, reward, value, pi, noise, noise_weight, seed, use_seed_dev):
        a = _lib.TreeArgs()
        a.B, a.num_simulations, a.sim = self.B, self.S, sim
        a.trees, a.s_tab, a.k_tab = _p(self.trees), _p(self.s_tab), _p(self.k_tab)
        a.discount, a.noise_weight, a.seed = self.discount, float(noise_weight), int(seed) & 0xFFFFFFFFFFFFFFFF
        a.reward, a.value, a.pi, a.noise = _p(reward), _p(value), _p(pi), _p(noise)
        a.leaf_parent, a.leaf_action, a.leaf_slot = _p(self.leaf_parent), _p(self.leaf_action), _p(self.leaf_slot)
        a.latent_store, a.dyn_in, a.latent_bytes = _p(self.latent_store), _p(self.dyn_in), self.latent_bytes
        a.out_value, a.out_visits, a.depth_hist = _p(self.out_value), _p(self.out_visits), _p(self.depth_hist)
        a.seed_dev = _p(self.seed_dev) if use_seed_dev else None
        return a

    def root(self, v_root, pi, noise, noise_weight, seed=0, use_seed_dev=False, stream=None):
        a = self._args(0, None, v_root, pi, noise, noise_weight, seed, use_seed_dev)
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mz_tree_root(C.byref(a), st))

    def step(self, sim, reward, value, pi, seed=0, use_seed_dev=False, stream=None):
        a = self._args(sim, reward, value, pi, None, 0.0, seed, use_seed_dev)
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().mz_tree_step(C.byref(a), st))


class _SearchPlan:
    """Everything preallocated for one (packed networks, B, S): trees, latent store, activation buffers,
    the root / per-simulation op programs and (optionally) the captured CUDA graph."""

    def __init__(self, nets: PackedNetworks, B, S, c1, c2, discount, noise_weight, use_graph):
        self.nets, self.B, self.S = nets, B, S
        dev = nets.device
        H, W = nets.latent_hw
        ch, hw = nets.latent_ch, H * W
        esz = 4 if nets.dt == F32 else 2
        self.tree = TreeBuffers(B, S, c1, c2, discount, dev, latent_bytes=hw * ch * esz)
        self.noise_weight = noise_weight
        t = self.tree
        store = t.latent_store.view(nets.dtype).view(B, t.nodes, hw, ch)
        dyn_in = t.dyn_in.view(nets.dtype).view(B, hw, ch)
        self.hidden = torch.empty((B, ch, H, W), dtype=torch.float32, device=dev)      # staged copy of the caller's root latents
        self.noise = torch.empty((B, 3), dtype=torch.float32, device=dev)
        self.pred_in = nets.buf(B, hw, ch)
        bufs = [nets.buf(B, hw, ch) for _ in range(3)]
        mid = nets.buf(B, hw, ch)
        f32 = nets.buf(B, hw, ch, torch.float32)
        self.reward = torch.zeros(B, dtype=torch.float32, device=dev)
        self.value = torch.zeros(B, dtype=torch.float32, device=dev)
        self.pi = torch.zeros((B, 3), dtype=torch.float32, device=dev)
        # root: NCHW fp32 latents -> channels-last pred_in + latent store slot 0, then the prediction network
        self.root_prog = Program(B)
        self.root_prog.add(op=OP_NCHW_IN, dtype=nets.dt, H=H, W=W, cin=ch, src=self.hidden, dst=self.pred_in, dst2=store,
                           dst2_slot=None, dst2_stride=t.nodes)
        self.root_prog.extend(nets.prediction_program(B, self.pred_in, bufs, mid, self.pi, self.value))
        # one simulation: dynamics on the gathered parent latents (+ scaled latent into the store at the leaf slot), prediction
        self.sim_prog = nets.dynamics_program(B, dyn_in, t.leaf_action, bufs, mid, f32, self.reward, self.pred_in, dst2=store,
                                              dst2_slot=t.leaf_slot, dst2_stride=t.nodes)
        self.sim_prog.extend(nets.prediction_program(B, self.pred_in, bufs, mid, self.pi, self.value))
        self.keep = [store, dyn_in, bufs, mid, f32]
        self.use_graph = use_graph
        self.graph = None
        self.conc = None              # (alpha, device tensor of concentrations) of the root noise, allocated on first use
        self.kernels_per_search = self.root_prog.n_kernels + 1 + S * (self.sim_prog.n_kernels + 1)

    def _enqueue(self, seed, use_seed_dev):
        t = self.tree
        self.root_prog.run()
        t.root(self.value, self.pi, self.noise, self.noise_weight, seed, use_seed_dev)
        for sim in range(self.S):
            self.sim_prog.run()
            t.step(sim, self.reward, self.value, self.pi, seed, use_seed_dev)

    def run(self, hidden, noise, seed):
        self.hidden.copy_(hidden, non_blocking=True)
        self.noise.copy_(noise, non_blocking=True)
        if not self.use_graph:
            self._enqueue(seed, False)
            return
        self.tree.seed_dev.fill_(int(np.uint64(seed & 0xFFFFFFFFFFFFFFFF).astype(np.int64)))
        if self.graph is None:
            self._enqueue(seed, True)                       # warm-up outside capture (lazy module / attribute init)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue(seed, True)
            self.graph = g
        self.graph.replay()


class MCTSSearchVec:
    def __init__(self, cfg, mu_zero, scalar_transforms=None):
        self.num_simulations = cfg["num_simulations"]
        self.actions = cfg["actions"]
        self.c1 = cfg["search"]["c1"]
        self.c2 = cfg["search"]["c2"]
        self.discount = cfg["search"]["discount_factor"]
        self.mu_zero = mu_zero
        self.scalar_transforms = scalar_transforms
        self.latent_resolution = cfg["latent_resolution"]
        self.dirchlet_alpha = 0.25          # (sic) mcts.py:21
        self.noise_weight = 0.175           # mcts.py:22
        if list(self.actions) != [0, 1, 2]:
            raise ValueError("the tree kernels are built for actions [0, 1, 2] (config.yaml:6)")
        s = cfg["search"]
        self.precision = s.get("precision", "f16")
        self.seed = int(s.get("seed", 0))
        self.use_graph = bool(s.get("use_graph", True))
        self.output_device = s.get("output_device", "cpu")
        self.device = torch.device(s.get("cuda_device", "cuda"))
        self._model_cfg = cfg.get("model", {})
        self._nets = None
        self._nets_external = False
        self._nets_key = None
        self._key_refs = None          # (module, its parameter / buffer tensors, their ids): see _weights_key
        self._plans = {}
        self.max_plans = int(s.get("max_plans", 3))
        self._calls = 0

    # ------------------------------------------------------------------ weights
    def _weights_key(self):
        """Identity of the module + the version counters of its parameters and buffers (load_state_dict and optimizer steps bump
        them in place, train_torch.py:137-138,361-367; :449-451 swaps the module).  Walking the module tree costs ~1.2 ms per call
        (542 tensors) -- 7 % of a 24-root search -- so the tensor list is cached per module object and only the version counters are
        read per call (0.07 ms); the full walk is repeated every 16th call in case a Parameter object itself was replaced."""
        m = self.mu_zero
        if isinstance(m, PackedNetworks):
            return ("packed", id(m))
        refs = self._key_refs
        if refs is None or refs[0] is not m or self._calls % 16 == 0:
            tensors = list(m.state_dict(keep_vars=True).values())
            self._key_refs = refs = (m, tensors, tuple(id(t) for t in tensors))
        return (id(m), refs[2]) + tuple(t._version for t in refs[1])

    def packed_networks(self) -> PackedNetworks:
        """(Re)pack when the module object or any parameter/buffer version changed
        (train_torch.py:137-138 load_state_dict, :449-451 swaps .mu_zero)."""
        key = self._weights_key()
        if key != self._nets_key:
            m = self.mu_zero
            if isinstance(m, PackedNetworks):
                self._nets, self._plans = m, {}
            else:
                fresh = PackedNetworks(m, self._model_cfg, self.precision, self.device)
                old = self._nets
                same = (old is not None and not self._nets_external and old.arenas.keys() == fresh.arenas.keys()
                        and all(old.arenas[k].shape == fresh.arenas[k].shape for k in fresh.arenas))
                if same:
                    # target refresh (train_torch.py:361-367): same architecture -> overwrite the weight arenas in place;
                    # every op program and captured CUDA graph keeps pointing at the same addresses
                    for k, a in fresh.arenas.items():
                        old.arenas[k].copy_(a)
                else:
                    self._nets, self._plans = fresh, {}
            self._nets_external = isinstance(m, PackedNetworks)
            self._nets_key = key
        return self._nets

    # ------------------------------------------------------------------ search
    def search(self, hidden_state: torch.Tensor, action_mask: torch.Tensor = None, training_iteration: int = 0, *,
               noise: torch.Tensor = None, seed: int = None):
        _lib.require_cuda()
        nets = self.packed_networks()
        B = hidden_state.shape[0]
        key = (B, self.num_simulations, float(self.noise_weight), self.c1, self.c2, self.discount)
        plan = self._plans.get(key)
        if plan is None:
            plan = _SearchPlan(nets, B, self.num_simulations, self.c1, self.c2, self.discount, float(self.noise_weight), self.use_graph)
            # keep the few most recent plans: train_torch.py alternates between n_parallel roots (acting) and 2 roots
            # (test rollout, :448-452) every iteration, and a plan is a graph capture plus its buffers
            self._plans[key] = plan
            while len(self._plans) > self.max_plans:
                self._plans.pop(next(iter(self._plans)))
        else:
            self._plans[key] = self._plans.pop(key)         # most recently used last
        if noise is None:                                   # Dirichlet(0.25 * ones(3)) per tree, mcts.py:114
            if plan.conc is None or plan.conc[0] != self.dirchlet_alpha:          # (alpha on the host, its device tensor)
                plan.conc = (self.dirchlet_alpha, torch.full((B, len(self.actions)), self.dirchlet_alpha, device=self.device))
            noise = torch._sample_dirichlet(plan.conc[1])
        if seed is None:
            seed = (self.seed * 0x9E3779B97F4A7C15 + self._calls * 0xD1B54A32D192ED03 + 1) & 0xFFFFFFFFFFFFFFFF
        self._calls += 1
        plan.run(hidden_state.to(self.device, torch.float32), noise.to(self.device, torch.float32), seed)
        value, visits = plan.tree.out_value, plan.tree.out_visits
        if self.output_device == "cpu":
            return value.cpu(), visits.cpu()
        return value.clone(), visits.clone()

    # kept for callers that use it directly (mcts.py:252-268)
    def _encode_action(self, action: torch.Tensor, resolution: tuple, n_actions: int):
        onehot = torch.nn.functional.one_hot(action, num_classes=n_actions).float()
        return onehot.view(action.shape[0], n_actions, 1, 1).expand(-1, -1, resolution[0], resolution[1])
