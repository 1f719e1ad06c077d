"""Host side of the MuZero networks on B200: packs the weights of a reference `MuZeroAgent`
(reference src/networks.py:245-350; state_dict layout in SURVEY.md Appendix D) into the layouts the
kernels of libmzb200.so want, and builds the op programs (struct mz_op, include/mzb200.h) that
evaluate the representation / dynamics / prediction networks.

    PackedNetworks(agent_or_state_dict, model_cfg, precision="bf16"|"f32")
        .representation(state)                 <-> MuZeroAgent.create_hidden_state_root  :271-280
        .dynamics(hidden, action_planes)       <-> MuZeroAgent.hidden_state_transition   :282-298
        .prediction(hidden)                    <-> MuZeroAgent.evaluate_state            :300-312
        .inverted_softmax_expectation(logits)  <-> ScalarTransforms (utils.py:74-81)

precision "f16" (the default of the search): fp16 weights and activations, tcgen05 tensor-core convolutions
(csrc/conv_tc.cu, conv_stack.cu), fp32 accumulation and epilogues, the residual stream of every ResidualBlock chain
carried as fp16 + an e4m3 correction plane (15 significant bits; csrc/tc_common.cuh) -- within 1e-3 of the fp32
reference (BASELINE.json north_star; profiles/emulate_precision.py); "bf16": the same path and speed with bf16 storage
(8-bit mantissa: 3-6e-3 of the reference); "f32": fp32 everything on CUDA cores (the 1e-5 parity path).
Eval-mode BatchNorm and the conv bias are folded: on the 16-bit paths the BatchNorm scale goes into the weights before
they are rounded and the epilogue adds a per-channel fp32 shift; the fp32 path keeps (scale, shift) in the epilogue.
The three one-hot action planes of the dynamics input become a per-action bias table (they are spatially constant,
so their convolution is a [3][20][256] lookup).
There is no CPU path: everything here needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn.functional as F

from .. import _lib

OP_CONV, OP_POOL2, OP_SCALE, OP_HEAD, OP_NCHW_IN, OP_NHWC_OUT = range(6)
F32, BF16, F16 = 0, 1, 2
FUSE_MAX_SAMPLES = int(os.environ.get("MZB_FUSE_MAX_SAMPLES", "1000000"))
# samples per slice of a trunk launch (all layers of a slice before the next slice, inside the one launch).  Measured on B200 at 4096 roots
# (profiles/README.md, round 2): one slice of 4096 3.82 ms per simulation step, two of 2048 4.06, four of 1024 5.16 -- fewer sample groups in
# flight leave the CTA pairs waiting on each other's layers, which costs more than the DRAM write-back of the 105 MB working set
STACK_SLICE = int(os.environ.get("MZB_STACK_SLICE", "4096"))
LO_STREAM = os.environ.get("MZB_NO_LO", "0") != "1"              # 16-bit residual streams carry an e4m3 correction plane
ACT = {"none": 0, "relu": 1, "leaky_relu": 2, "silu": 3, "gelu": 4}   # utils.py:99-108


class MzOp(C.Structure):
    """mirror of struct mz_op (include/mzb200.h)"""
    _fields_ = [("op", C.c_int32), ("dtype", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("cin", C.c_int32),
                ("cout", C.c_int32), ("ksize", C.c_int32), ("act", C.c_int32), ("use_tc", C.c_int32), ("nout", C.c_int32),
                ("head_mode", C.c_int32), ("w_layout", C.c_int32),
                ("src", C.c_void_p), ("dst", C.c_void_p), ("res", C.c_void_p), ("dst_f32", C.c_void_p), ("w", C.c_void_p),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("act_bias", C.c_void_p), ("act_idx", C.c_void_p),
                ("dst2", C.c_void_p), ("dst2_slot", C.c_void_p), ("dst2_stride", C.c_int64),
                ("out", C.c_void_p), ("out_logits", C.c_void_p), ("res_lo", C.c_void_p), ("dst_lo", C.c_void_p), ("res_f32", C.c_void_p), ("bn_partial", C.c_void_p)]


def _p(t):
    return None if t is None else t.data_ptr()


def _stackable(o) -> bool:
    """3x3 / 1x1 256->256 tensor-core convolution on the 4x5 latent with tile-contiguous, scale-folded weights (csrc/conv_stack.cu)."""
    return (o.op == OP_CONV and o.dtype in (BF16, F16) and o.use_tc == 1 and o.w_layout == 1 and o.ksize in (1, 3) and o.cin == 256
            and o.cout == 256 and o.H == 4 and o.W == 5 and not o.scale)


def _head_mma_ok(o) -> bool:
    """mirror of csrc/nets.cu head_mma_ok: the heads that run on the warp-MMA kernel (two consecutive ones share a launch)"""
    feat = o.H * o.W * o.cin
    cstride = o.cout if o.cout > 0 else o.cin
    return (o.op == OP_HEAD and o.dtype in (BF16, F16) and o.cin % 32 == 0 and cstride % 8 == 0 and cstride >= o.cin and 1 <= o.nout <= 16
            and feat % 1280 == 0 and feat // 256 <= 20 and os.environ.get("MZB_HEAD_SIMT", "0") in ("", "0"))


def lat_max_samples() -> int:
    """Batches up to this size run their trunks in latency mode (csrc/conv_lat.cu: a single wave of 16 x ceil(n/3) work
    items); larger ones on the tcgen05 trunk (csrc/conv_stack.cu).  MZB_LAT_MAX_SAMPLES overrides (0 = never)."""
    e = os.environ.get("MZB_LAT_MAX_SAMPLES")
    return int(e) if e is not None else int(_lib.lib().mz_lat_max_samples())


def _lat_stackable(o) -> bool:
    """what the latency-mode trunk (csrc/conv_lat.cu) takes: 3x3 or 1x1 256->256 convolutions on the 4x5 latent, and a final pair
    of 256->128 ones that read the same input (the policy and value heads' ConvBlocks)."""
    return (o.op == OP_CONV and o.dtype in (BF16, F16) and o.use_tc == 1 and o.w_layout == 1 and o.ksize in (1, 3) and o.cin == 256
            and o.cout in (128, 256) and o.H == 4 and o.W == 5)


class _Stack:
    """A run of stackable convolutions executed by one persistent launch (mz_stack_run, or mz_lat_run for small batches)."""

    def __init__(self, ops, n, device, lat_max=None, n_tail=0):
        """ops: the convolutions, followed (latency mode only) by n_tail head / scale ops that run at the end of the same launch."""
        L = _lib.lib()
        self.n, self.nlayers = n, len(ops)
        self.act_idx = next((o.act_idx for o in ops if o.act_idx), None)
        self.dtype = ops[0].dtype
        arr = (MzOp * self.nlayers)(*ops)
        convs = ops[:len(ops) - n_tail]
        self.lat = n <= (lat_max_samples() if lat_max is None else lat_max) and len(convs) <= L.mz_lat_max_layers() and all(_lat_stackable(o) for o in convs)
        assert self.lat or n_tail == 0
        if self.lat:
            self.ok = True
            lb = L.mz_lat_layer_bytes()
            raw = (C.c_uint8 * (self.nlayers * lb + 64))()
            host = (C.addressof(raw) + 63) & ~63
            rc = L.mz_lat_build(arr, self.nlayers, host, self.nlayers * lb)
            if rc < 0:
                _lib.check(rc)
            self.flags = rc                              # bit 0: the last two convolutions are the halves of one layer; bits 1-2: tail ops
            self.blob = torch.frombuffer((C.c_uint8 * (self.nlayers * lb)).from_address(host), dtype=torch.uint8).clone().to(device)
            self.chunks = [(0, n)]
            self.done = torch.zeros((L.mz_lat_scratch_bytes(n) + 3) // 4, dtype=torch.int32, device=device)   # zeroed once; the launches keep their epoch in it
            return
        bufs = []
        for o in ops:
            for ptr in (o.src, o.dst, o.res):
                if ptr and ptr not in bufs:
                    bufs.append(ptr)
        self.ok = len(bufs) <= 5                         # MZ_STACK_MAX_BUFS
        if not self.ok:
            return
        lb = L.mz_stack_layer_bytes()
        raw = (C.c_uint8 * (self.nlayers * lb + 64))()
        host = (C.addressof(raw) + 63) & ~63
        self.bufs = (C.c_void_p * len(bufs))(*bufs)
        self.nbufs = len(bufs)
        _lib.check(L.mz_stack_build(arr, self.nlayers, host, self.nlayers * lb, self.bufs, self.nbufs))
        blob = torch.frombuffer((C.c_uint8 * (self.nlayers * lb)).from_address(host), dtype=torch.uint8).clone()
        self.blob = blob.to(device)
        # one launch; inside it the samples are walked in slices of ~STACK_SLICE (whole 256-sample pair tiles, equal sizes)
        nsl = 1 if n <= STACK_SLICE * 5 // 4 else (n + STACK_SLICE - 1) // STACK_SLICE
        self.slice = 0 if nsl == 1 else ((n + nsl - 1) // nsl + 255) // 256 * 256
        self.chunks = [(0, n)]
        self.done = torch.zeros((L.mz_stack_scratch_bytes(self.nlayers, n) + 3) // 4, dtype=torch.int32, device=device)   # zeroed once; the launches keep their epoch in it

    def run(self, st):
        L = _lib.lib()
        if self.lat:
            _lib.check(L.mz_lat_run(self.blob.data_ptr(), self.nlayers, self.flags, self.n, self.act_idx, self.done.data_ptr(), self.dtype, st))
            return
        _lib.check(L.mz_stack_run(self.blob.data_ptr(), self.nlayers, 0, self.n, self.slice, self.bufs, self.nbufs, self.act_idx,
                                  self.done.data_ptr(), self.dtype, st))


class Program:
    """A list of mz_op records with every pointer resolved; run() enqueues it on the current stream.  With
    fuse=True, runs of two or more stackable convolutions become one persistent launch each."""

    def __init__(self, n: int, fuse: bool = False, lat_max: int | None = None):
        self.n = n
        self.ops = []
        self.keep = []            # tensors the ops point into
        self.fuse = fuse
        self.lat_max = lat_max    # None: lat_max_samples(); 0: tcgen05 trunk at every batch size
        self._segs = None

    def add(self, **kw):
        op = MzOp()
        for k, v in kw.items():
            if isinstance(v, torch.Tensor):
                self.keep.append(v)
                v = v.data_ptr()
            setattr(op, k, v)
        self.ops.append(op)
        self._segs = None
        return op

    def extend(self, other: "Program"):
        self.ops += other.ops
        self.keep += other.keep
        self.fuse = self.fuse or other.fuse
        self.lat_max = other.lat_max if self.lat_max is None else self.lat_max
        self._segs = None

    def _build(self):
        segs, plain, i = [], [], 0
        device = next((t.device for t in self.keep if t.is_cuda), torch.device("cuda"))

        def flush():
            if plain:
                segs.append(("ops", (MzOp * len(plain))(*plain), len(plain)))
                plain.clear()

        while i < len(self.ops):
            j = i
            # measured (1 B200, 50-simulation searches, profiles/prof_trunk.py): with the dependency scout warp and the per-layer
            # rotation of the tile assignment the persistent trunk launch beats one launch per layer at every batch size
            # (4096: 3.94 -> 3.67 ms per simulation step), run over <= ~4096-sample slices so that the working set stays in L2
            fuse = self.fuse and self.n <= FUSE_MAX_SAMPLES
            lat = self.n <= (lat_max_samples() if self.lat_max is None else self.lat_max)
            if lat:
                # latency mode also takes 1x1 256->256 layers (the reward head's ConvBlock) and needs a chain: each layer reads
                # the previous one's output
                max_layers = _lib.lib().mz_lat_max_layers()
                ops = self.ops
                while fuse and j < len(ops) and j - i < max_layers and _lat_stackable(ops[j]) and (j == i or ops[j].src == ops[j - 1].dst):
                    if ops[j].cout == 256:
                        j += 1
                        continue
                    # 256 -> 128: only as the final pair of a run, both halves reading the previous layer's output
                    if (j > i and j + 1 < len(ops) and j - i + 2 <= max_layers and _lat_stackable(ops[j + 1]) and ops[j + 1].cout == 128
                            and ops[j + 1].src == ops[j].src and ops[j + 1].dst != ops[j].dst and not ops[j].res and not ops[j + 1].res):
                        j += 2
                    break
            else:
                while fuse and j < len(self.ops) and _stackable(self.ops[j]):
                    j += 1
            n_tail = 0
            if lat and j - i >= 2:
                # the heads (Flatten + Linear + softmax / expectation) and _scale_state that follow run at the end of the same launch
                while (n_tail < 2 and j + n_tail < len(self.ops) and self.ops[j + n_tail].dtype == self.ops[i].dtype
                       and (self.ops[j + n_tail].op == OP_HEAD and self.ops[j + n_tail].nout in (3, 11)
                            or self.ops[j + n_tail].op == OP_SCALE and self.ops[j + n_tail].H * self.ops[j + n_tail].W * self.ops[j + n_tail].cin == 5120)):
                    n_tail += 1
            stack = _Stack(self.ops[i:j + n_tail], self.n, device, self.lat_max, n_tail) if j - i >= 2 else None
            j += n_tail if stack is not None and stack.ok else 0
            if stack is not None and stack.ok:
                flush()
                segs.append(("stack", stack, j - i))          # j - i: ops covered by the launch
                i = j
            else:
                plain.append(self.ops[i])
                i += 1
        flush()
        self._segs = segs

    def run(self, stream=None):
        if self._segs is None:
            self._build()
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        for kind, item, cnt in self._segs:
            if kind == "ops":
                _lib.check(_lib.lib().mz_run(item, cnt, self.n, st))
            else:
                item.run(st)

    @property
    def n_kernels(self) -> int:
        """kernel launches of one run(): one per plain op record -- except that two consecutive 16-bit heads share a launch (csrc/nets.cu
        head_mma_kernel, the policy + value pair) -- and one per fused trunk."""
        if self._segs is None:
            self._build()
        total = 0
        for kind, item, cnt in self._segs:
            if kind != "ops":
                total += len(item.chunks)
                continue
            i = 0
            while i < cnt:
                if _head_mma_ok(item[i]) and i + 1 < cnt and _head_mma_ok(item[i + 1]) and item[i + 1].dtype == item[i].dtype:
                    i += 1
                total += 1
                i += 1
        return total


DEFAULT_MODEL_CFG = {  # config.yaml:27-50
    "num_supports": 11, "supports_min": -5, "supports_max": 5, "latent_channels": [128, 256], "state_history_length": 32,
    "latent_resolution": [4, 5],
    "representation_network": {"num_res_blocks": [2, 3, 3], "activation": "relu"},
    "dynamics_network": {"num_res_blocks": 14, "num_actions": 3, "activation": "relu"},
    "prediction_network": {"num_res_blocks": 14, "num_actions": 3, "activation": "relu"},
}


def random_state_dict(cfg: dict | None = None, seed: int = 0, bn_jitter: float = 0.0) -> dict:
    """Random-init weights with the key names and shapes of MuZeroAgent.state_dict():
    uniform(+-1/sqrt(fan_in)) convs / linears, BatchNorm at its defaults (optionally jittered).  For
    benchmarks and smoke tests, where no checkpoint is available."""
    cfg = cfg or DEFAULT_MODEL_CFG
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(key, cout, cin, k):
        bound = 1.0 / (cin * k * k) ** 0.5
        sd[key + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        sd[key + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def bn(key, c):
        j = bn_jitter
        sd[key + ".weight"] = 1 + j * (torch.rand(c, generator=g) - 0.5)
        sd[key + ".bias"] = j * (torch.rand(c, generator=g) - 0.5)
        sd[key + ".running_mean"] = j * (torch.rand(c, generator=g) - 0.5)
        sd[key + ".running_var"] = 1 + j * (torch.rand(c, generator=g) - 0.5)

    def res(key, c):
        conv(key + ".conv1", c, c, 3); bn(key + ".bn1", c); conv(key + ".conv2", c, c, 3); bn(key + ".bn2", c)

    def lin(key, nout, nin):
        bound = 1.0 / nin ** 0.5
        sd[key + ".weight"] = (torch.rand(nout, nin, generator=g) * 2 - 1) * bound
        sd[key + ".bias"] = (torch.rand(nout, generator=g) * 2 - 1) * bound

    c0, c1 = cfg["latent_channels"]
    n0, n1, n2 = cfg["representation_network"]["num_res_blocks"]
    hw = cfg["latent_resolution"][0] * cfg["latent_resolution"][1]
    i = 0
    conv(f"rep_net.blocks.{i}", c0, cfg["state_history_length"] * 2, 3); i += 1
    for _ in range(n0):
        res(f"rep_net.blocks.{i}", c0); i += 1
    conv(f"rep_net.blocks.{i}", c1, c0, 3); i += 1
    for _ in range(n1):
        res(f"rep_net.blocks.{i}", c1); i += 1
    i += 1                                                   # AvgPool2d
    for _ in range(n2):
        res(f"rep_net.blocks.{i}", c1); i += 1
    na = cfg["dynamics_network"]["num_actions"]
    conv("dyn_net.conv_block.conv", c1, c1 + na, 3); bn("dyn_net.conv_block.bn", c1)
    for b in range(cfg["dynamics_network"]["num_res_blocks"]):
        res(f"dyn_net.res_blocks.{b}", c1)
    conv("dyn_net.reward_head.0.conv", c1, c1, 1); bn("dyn_net.reward_head.0.bn", c1)
    lin("dyn_net.reward_head.2", cfg["num_supports"], c1 * hw)
    for b in range(cfg["prediction_network"]["num_res_blocks"]):
        res(f"pred_net.res_blocks.{b}", c1)
    conv("pred_net.policy_head.0.conv", c1 // 2, c1, 3); bn("pred_net.policy_head.0.bn", c1 // 2)
    lin("pred_net.policy_head.2", cfg["prediction_network"]["num_actions"], c1 // 2 * hw)
    conv("pred_net.value_head.0.conv", c1 // 2, c1, 1); bn("pred_net.value_head.0.bn", c1 // 2)
    lin("pred_net.value_head.2", cfg["num_supports"], c1 // 2 * hw)
    return sd


class _Conv:
    """One packed convolution: w [cout][k*k*cin] (tap-major, then input channel), fp32 scale/shift."""

    def __init__(self, w, scale, shift, ksize, act, act_bias=None):
        self.w, self.scale, self.shift, self.ksize, self.act, self.act_bias = w, scale, shift, ksize, act, act_bias
        self.cout = w.shape[0]
        self.cin = w.shape[1] // (ksize * ksize)
        self.w_layout = 0


class _Lin:
    """One packed Linear head: fp32 w [nout][H*W*C] in channels-last flatten order, fp32 bias."""

    def __init__(self, w, b, nout):
        self.w, self.b, self.nout = w, b, nout


class PackedNetworks:
    def __init__(self, agent, model_cfg: dict | None = None, precision: str = "bf16", device="cuda", use_tc: bool | None = None):
        _lib.require_cuda()
        if precision not in ("bf16", "f16", "f32"):
            raise ValueError("precision must be 'bf16', 'f16' or 'f32'")
        sd = agent if isinstance(agent, dict) else agent.state_dict()
        cfg = model_cfg or getattr(agent, "cfg", None) or {}
        self.device = torch.device(device)
        self.precision = precision
        self.dtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[precision]
        self.dt = {"bf16": BF16, "f16": F16, "f32": F32}[precision]
        half = precision in ("bf16", "f16")
        self.use_tc = half if use_tc is None else bool(use_tc and half)
        self.fold_scale = half
        self.lo_stream = half and LO_STREAM
        self.fuse_stacks = self.use_tc and os.environ.get("MZB_NO_STACK", "0") != "1"   # whole trunks in one persistent launch
        self.lat_max = None        # None: trunks of batches <= lat_max_samples() run in latency mode; 0: always the tcgen05 trunk
        self.num_supports = int(cfg.get("num_supports", 11))
        self.supports_min, self.supports_max = cfg.get("supports_min", -5), cfg.get("supports_max", 5)
        if self.supports_min != -self.supports_max or self.supports_max - self.supports_min != self.num_supports - 1:
            raise ValueError("the fused head kernel assumes unit-spaced supports centred on 0, e.g. linspace(-5, 5, 11) (config.yaml:30-32)")
        acts = {k: ACT[cfg.get(k, {}).get("activation", "relu")] for k in ("representation_network", "dynamics_network", "prediction_network")}
        sd = {k: v.detach().to("cpu", torch.float64) for k, v in sd.items() if v.dtype.is_floating_point}
        self._pack(sd, acts)

    # ------------------------------------------------------------------ packing
    def _dev(self, t, dtype):
        return t.to(self.device, dtype).contiguous()

    def _conv(self, sd, conv_key, bn_key, act, cin_used=None):
        w = sd[conv_key + ".weight"]                       # (cout, cin, k, k)
        b = sd[conv_key + ".bias"]
        cout, cin, k, _ = w.shape
        if bn_key is not None:                             # eval-mode BN (eps 1e-5) folded: y = conv*a + (b - mean)*a + beta
            a = sd[bn_key + ".weight"] / torch.sqrt(sd[bn_key + ".running_var"] + 1e-5)
            shift = (b - sd[bn_key + ".running_mean"]) * a + sd[bn_key + ".bias"]
        else:
            a, shift = torch.ones(cout, dtype=torch.float64), b
        if self.fold_scale:                                # 16-bit paths: the scale goes into the weights before they are rounded
            w = w * a.view(-1, 1, 1, 1)
            a = None
        act_bias = None
        if cin_used is not None and cin_used < cin:        # dynamics conv_block: channels >= cin_used are the action planes
            H, W = self.latent_hw
            planes = []
            for ch in range(cin_used, cin):
                o = F.conv2d(torch.ones(1, 1, H, W, dtype=torch.float64), w[:, ch:ch + 1], padding=k // 2)   # (1,cout,H,W)
                planes.append(o[0].permute(1, 2, 0).reshape(H * W, cout))
            act_bias = self._dev(torch.stack(planes), torch.float32)                                          # [3][HW][cout]
            w = w[:, :cin_used]
        wp = w.permute(0, 2, 3, 1).reshape(cout, -1)       # [cout][(ky*k+kx)*cin + c]
        cv = _Conv(self._dev(wp, self.dtype), None if a is None else self._dev(a, torch.float32), self._dev(shift, torch.float32), k, act, act_bias)
        if self.use_tc and cv.cin % 64 == 0:               # tile-contiguous copy for the TMA loads: [tap][cin/64][cout][64]
            cv.w = self._dev(wp.reshape(cout, k * k, cv.cin // 64, 64).permute(1, 2, 0, 3), self.dtype)
            cv.w_layout = 1
        return cv

    def _pv_conv(self, sd, act):
        """Policy (3x3 256->128) and value (1x1 256->128) head ConvBlocks (networks.py:200-218) as ONE 3x3 256->256 convolution: output
        channels 0-127 = the policy ConvBlock, 128-255 = the value ConvBlock with its 1x1 weights in the centre tap (the other eight
        taps are exact zeros, so the sums are the 1x1 convolution's).  The tensor-core trunk runs it as its last layer."""
        pk, vk = "pred_net.policy_head.0", "pred_net.value_head.0"
        wp_, wv_ = sd[pk + ".conv.weight"], sd[vk + ".conv.weight"]
        if not (self.use_tc and wp_.shape[2] == 3 and wv_.shape[2] == 1 and wp_.shape[1] == wv_.shape[1] == self.latent_ch
                and wp_.shape[0] + wv_.shape[0] == self.latent_ch):
            return None
        wv3 = torch.zeros(wv_.shape[0], wv_.shape[1], 3, 3, dtype=wv_.dtype)
        wv3[:, :, 1, 1] = wv_[:, :, 0, 0]
        both = {"c.weight": torch.cat([wp_, wv3]), "c.bias": torch.cat([sd[pk + ".conv.bias"], sd[vk + ".conv.bias"]])}
        for k in ("weight", "bias", "running_mean", "running_var"):
            both["b." + k] = torch.cat([sd[pk + ".bn." + k], sd[vk + ".bn." + k]])
        return self._conv(both, "c", "b", act)

    def _res(self, sd, prefix, act):
        return (self._conv(sd, prefix + ".conv1", prefix + ".bn1", act), self._conv(sd, prefix + ".conv2", prefix + ".bn2", act))

    def _linear(self, sd, key, C_, HW):
        w = sd[key + ".weight"]                            # (nout, C*HW), column = c*HW + p  (Flatten of NCHW)
        nout = w.shape[0]
        wp = w.reshape(nout, C_, HW).permute(0, 2, 1).reshape(nout, HW * C_)   # column = p*C + c (channels-last)
        return _Lin(self._dev(wp, torch.float32), self._dev(sd[key + ".bias"], torch.float32), nout)

    def _pack(self, sd, acts):
        rep_keys = sorted({int(k.split(".")[2]) for k in sd if k.startswith("rep_net.blocks.")})
        self.latent_ch = sd["dyn_net.reward_head.0.conv.weight"].shape[1]
        hw = sd["dyn_net.reward_head.2.weight"].shape[1] // self.latent_ch
        self.latent_hw = {20: (4, 5)}.get(hw, None)
        if self.latent_hw is None:
            raise ValueError(f"unsupported latent resolution ({hw} pixels)")
        # representation: plain convs (bias only) and residual blocks; pools sit where block indices are missing
        self.rep = []
        a = acts["representation_network"]
        for i in range(max(rep_keys) + 2):
            if f"rep_net.blocks.{i}.weight" in sd:
                self.rep.append(("conv", self._conv(sd, f"rep_net.blocks.{i}", None, ACT["none"])))
            elif f"rep_net.blocks.{i}.conv1.weight" in sd:
                self.rep.append(("res", self._res(sd, f"rep_net.blocks.{i}", a)))
            else:
                self.rep.append(("pool", None))            # nn.AvgPool2d has no parameters (blocks.7 / blocks.11)
        self.rep_cin = self.rep[0][1].cin
        a = acts["dynamics_network"]
        self.dyn_first = self._conv(sd, "dyn_net.conv_block.conv", "dyn_net.conv_block.bn", a, cin_used=self.latent_ch)
        n_dyn = len({k.split(".")[2] for k in sd if k.startswith("dyn_net.res_blocks.")})
        self.dyn_res = [self._res(sd, f"dyn_net.res_blocks.{i}", a) for i in range(n_dyn)]
        self.reward_conv = self._conv(sd, "dyn_net.reward_head.0.conv", "dyn_net.reward_head.0.bn", a)
        self.reward_lin = self._linear(sd, "dyn_net.reward_head.2", self.reward_conv.cout, hw)
        a = acts["prediction_network"]
        n_pred = len({k.split(".")[2] for k in sd if k.startswith("pred_net.res_blocks.")})
        self.pred_res = [self._res(sd, f"pred_net.res_blocks.{i}", a) for i in range(n_pred)]
        self.policy_conv = self._conv(sd, "pred_net.policy_head.0.conv", "pred_net.policy_head.0.bn", a)
        self.policy_lin = self._linear(sd, "pred_net.policy_head.2", self.policy_conv.cout, hw)
        self.value_conv = self._conv(sd, "pred_net.value_head.0.conv", "pred_net.value_head.0.bn", a)
        self.value_lin = self._linear(sd, "pred_net.value_head.2", self.value_conv.cout, hw)
        self.num_actions = self.policy_lin.nout
        self.pv_conv = self._pv_conv(sd, a)
        self._consolidate()

    def holders(self):
        """(object, attribute) of every packed device tensor, in a deterministic order."""
        out = []

        def conv(c):
            out.extend((c, a) for a in ("w", "scale", "shift", "act_bias") if getattr(c, a) is not None)

        for kind, item in self.rep:
            if kind == "conv":
                conv(item)
            elif kind == "res":
                conv(item[0]); conv(item[1])
        conv(self.dyn_first)
        for a, b in self.dyn_res + self.pred_res:
            conv(a); conv(b)
        for c, lin in ((self.reward_conv, self.reward_lin), (self.policy_conv, self.policy_lin), (self.value_conv, self.value_lin)):
            conv(c)
            out.extend([(lin, "w"), (lin, "b")])
        if self.pv_conv is not None:
            conv(self.pv_conv)
        return out

    def _consolidate(self):
        """Move every packed tensor into one flat arena per dtype (256-byte aligned slices) and replace it by a
        view: a target-network refresh is then ONE NCCL broadcast per arena, with no gather/scatter copies."""
        self.arenas = {}
        by_dtype = {}
        for obj, attr in self.holders():
            by_dtype.setdefault(getattr(obj, attr).dtype, []).append((obj, attr))
        for dtype, items in by_dtype.items():
            esz = torch.empty((), dtype=dtype).element_size()
            align = 256 // esz
            offs, total = [], 0
            for obj, attr in items:
                offs.append(total)
                total += (getattr(obj, attr).numel() + align - 1) // align * align
            arena = torch.zeros(total, dtype=dtype, device=self.device)
            for (obj, attr), off in zip(items, offs):
                t = getattr(obj, attr)
                view = arena[off:off + t.numel()].view(t.shape)
                view.copy_(t)
                setattr(obj, attr, view)
            self.arenas[dtype] = arena

    # ------------------------------------------------------------------ program building
    def buf(self, n, hw, c, dtype=None):
        return torch.empty((n, hw, c), dtype=dtype or self.dtype, device=self.device)

    def _add_conv(self, prog, cv: _Conv, H, W, src, dst, res=None, dst_f32=None, act_idx=None, res_lo=None, dst_lo=None, res_f32=None):
        prog.add(op=OP_CONV, dtype=self.dt, H=H, W=W, cin=cv.cin, cout=cv.cout, ksize=cv.ksize, act=cv.act,
                 use_tc=int(self.use_tc), w_layout=cv.w_layout, src=src, dst=dst, res=res, dst_f32=dst_f32, w=cv.w, scale=cv.scale, shift=cv.shift,
                 act_bias=cv.act_bias if act_idx is not None else None, act_idx=act_idx, res_lo=res_lo, dst_lo=dst_lo, res_f32=res_f32)

    def lo_plane(self, n, H, W, c):
        """Correction plane of a 16-bit residual stream [n][H*W][c] (csrc/tc_common.cuh: split2 / lo2), or None on the fp32 path."""
        if not self.lo_stream:
            return None
        nbytes = int(_lib.lib().mz_conv_lo_bytes(n, H, W, c, 3))
        return torch.zeros(nbytes, dtype=torch.uint8, device=self.device)

    def _add_res_blocks(self, prog, blocks, H, W, bufs, cur, last_f32=None, mid=None, lo=None, lo_valid=False, keep_lo=False):
        """bufs: same-shaped activation buffers; cur: index of the one holding the input.  Each block writes its
        output IN PLACE over its input: conv2's epilogue reads the residual element and then writes the result to
        the same address from the same thread, and no tile reads the block input during conv2 (its operand is the
        mid buffer).  lo: the stream's correction plane (lo_valid: it holds the correction of the input); the last block
        keeps it only if keep_lo (something after it continues the stream).
        Returns the index holding the output (= cur)."""
        mid = (cur + 1) % len(bufs) if mid is None else mid
        for i, (c1, c2) in enumerate(blocks):
            last = i == len(blocks) - 1
            self._add_conv(prog, c1, H, W, bufs[cur], bufs[mid])
            self._add_conv(prog, c2, H, W, bufs[mid], bufs[cur], res=bufs[cur], dst_f32=last_f32 if last else None,
                           res_lo=lo if lo_valid else None, dst_lo=lo if (lo is not None and (keep_lo or not last)) else None)
            lo_valid = lo is not None
        return cur

    def _add_head(self, prog, conv, lin, H, W, src, mid, mode, out, out_logits=None):
        self._add_conv(prog, conv, H, W, src, mid)
        w, b, nout = lin.w, lin.b, lin.nout
        prog.add(op=OP_HEAD, dtype=self.dt, H=H, W=W, cin=conv.cout, nout=nout, head_mode=mode, src=mid, w=w, shift=b,
                 out=out, out_logits=out_logits)

    def _lat(self, n) -> bool:
        """does a batch of n samples run its trunks in latency mode (csrc/conv_lat.cu)?"""
        return self.use_tc and self.fuse_stacks and n <= (lat_max_samples() if self.lat_max is None else self.lat_max)

    def prediction_program(self, n, src, bufs, mid, pi, value, policy_logits=None, value_logits=None, value_mode=1, pi_mode=2, src_f32=None):
        """14 residual blocks + policy head + value head (networks.py:225-241) on `src` [n][20][256].  src_f32: the same latents as
        float32 [n][20][256] when the caller has them in full precision (the first block's residual is then exact; the search's
        latents come from the 16-bit store and are exact as they are)."""
        H, W = self.latent_hw
        prog = Program(n, self.fuse_stacks, self.lat_max)
        lo = self.lo_plane(n, H, W, self.latent_ch)
        if lo is not None:
            prog.keep.append(lo)
        c1, c2 = self.pred_res[0]                                   # first block reads src directly (src, bufs[0], bufs[1]: three buffers)
        self._add_conv(prog, c1, H, W, src, bufs[0])
        f32_res = src_f32 if (src_f32 is not None and self.dt != F32) else None
        self._add_conv(prog, c2, H, W, bufs[0], bufs[1], res=None if f32_res is not None else src, res_f32=f32_res,
                       dst_lo=lo if len(self.pred_res) > 1 else None)   # src from the 16-bit latent store is exact: no correction plane
        cur = self._add_res_blocks(prog, self.pred_res[1:], H, W, bufs, 1, mid=0, lo=lo, lo_valid=lo is not None)
        pc, vc = self.policy_conv, self.value_conv
        if self.pv_conv is not None and not self._lat(n) and mid.shape[-1] == self.pv_conv.cout:
            # tensor-core trunk: both head ConvBlocks as ONE more 3x3 256->256 layer of the trunk launch (the value head's 1x1 weights sit in
            # the centre tap of output channels 128-255); the two Linear heads read the halves of its 256-channel rows
            self._add_conv(prog, self.pv_conv, H, W, bufs[cur], mid)
            mid_v = mid.view(-1)[pc.cout:]
            for lin, conv, src_, mode, out, logits in ((self.policy_lin, pc, mid, pi_mode, pi, policy_logits), (self.value_lin, vc, mid_v, value_mode, value, value_logits)):
                prog.add(op=OP_HEAD, dtype=self.dt, H=H, W=W, cin=conv.cout, cout=self.pv_conv.cout, nout=lin.nout, head_mode=mode, src=src_, w=lin.w, shift=lin.b,
                         out=out, out_logits=logits)
            return prog
        # both head ConvBlocks first (they read the same trunk output and write the two halves of `mid`), then the two Linear heads:
        # in latency mode the pair rides as the last, split layer of the trunk launch
        half = n * H * W * pc.cout
        mid_p = mid.view(-1)[:half].view(n, H * W, pc.cout)
        if pc.cout + vc.cout <= mid.shape[-1]:
            mid_v = mid.view(-1)[half:half + n * H * W * vc.cout].view(n, H * W, vc.cout)
        else:                                                      # head ConvBlocks wider than half the latent: a buffer of its own
            mid_v = self.buf(n, H * W, vc.cout)
            prog.keep.append(mid_v)
        self._add_conv(prog, pc, H, W, bufs[cur], mid_p)
        self._add_conv(prog, vc, H, W, bufs[cur], mid_v)
        for lin, conv, src_, mode, out, logits in ((self.policy_lin, pc, mid_p, pi_mode, pi, policy_logits),
                                                   (self.value_lin, vc, mid_v, value_mode, value, value_logits)):
            prog.add(op=OP_HEAD, dtype=self.dt, H=H, W=W, cin=conv.cout, nout=lin.nout, head_mode=mode, src=src_, w=lin.w, shift=lin.b,
                     out=out, out_logits=logits)
        return prog

    def dynamics_program(self, n, src, act_idx, bufs, mid, f32, reward, dst, dst2=None, dst2_slot=None, dst2_stride=0,
                         reward_logits=None, reward_mode=1):
        """ConvBlock(259->256) + 14 residual blocks + reward head + _scale_state (networks.py:151-167,
        282-298).  src [n][20][256] parent latents, act_idx int32 [n]; scaled latent -> dst (and dst2)."""
        H, W = self.latent_hw
        prog = Program(n, self.fuse_stacks, self.lat_max)
        lo = self.lo_plane(n, H, W, self.latent_ch)
        if lo is not None:
            prog.keep.append(lo)
        self._add_conv(prog, self.dyn_first, H, W, src, bufs[0], act_idx=act_idx, dst_lo=lo if self.dyn_res else None)
        cur = self._add_res_blocks(prog, self.dyn_res, H, W, bufs, 0, last_f32=f32, lo=lo, lo_valid=lo is not None)
        self._add_head(prog, self.reward_conv, self.reward_lin, H, W, bufs[cur], mid, reward_mode, reward, reward_logits)
        prog.add(op=OP_SCALE, dtype=self.dt, H=H, W=W, cin=self.latent_ch, src=f32, dst=dst, dst2=dst2, dst2_slot=dst2_slot,
                 dst2_stride=dst2_stride)
        return prog

    def representation_program(self, n, x_nchw, out_nchw, x_cl=None):
        """RepresentationNetwork.forward + _scale_state (networks.py:94-99, 271-280): float32 NCHW in/out, or a
        channels-last input buffer x_cl [n][320][64] of the activation dtype (then x_nchw is ignored)."""
        H, W = 16, 20
        prog = Program(n)
        cmax = max(cv.cout for kind, cv in self.rep if kind == "conv")
        bufs = [self.buf(n, H * W, cmax) for _ in range(3)]
        f32 = self.buf(n, self.latent_hw[0] * self.latent_hw[1], cmax, torch.float32)
        if x_cl is None:
            xin = self.buf(n, H * W, self.rep_cin)
            prog.add(op=OP_NCHW_IN, dtype=self.dt, H=H, W=W, cin=self.rep_cin, src=x_nchw, dst=xin)
        else:
            xin = x_cl
        src, cur = xin, None
        n_pool = sum(1 for kind, _ in self.rep if kind == "pool")
        pools = 0
        # the residual stream's correction plane (one buffer: it follows the stream; a plane is only ever exchanged between two
        # convolutions of the same shape, and a pool / plain convolution reads the 16-bit stream alone)
        lo = self.lo_plane(n, H, W, cmax)
        if lo is not None:
            prog.keep.append(lo)
        lo_valid = False
        for idx, (kind, item) in enumerate(self.rep):
            nxt_kind = self.rep[idx + 1][0] if idx + 1 < len(self.rep) else None
            if kind == "conv":
                nxt = 0 if cur is None else (cur + 1) % 3
                self._add_conv(prog, item, H, W, src, bufs[nxt], dst_lo=lo if nxt_kind == "res" else None)
                lo_valid = lo is not None and nxt_kind == "res"
                cur, src = nxt, bufs[nxt]
                ch = item.cout
            elif kind == "res":
                cur = self._add_res_blocks(prog, [item], H, W, bufs, cur, lo=lo, lo_valid=lo_valid, keep_lo=nxt_kind == "res")
                lo_valid = lo is not None and nxt_kind == "res"
                src = bufs[cur]
            else:
                lo_valid = False
                pools += 1
                nxt = (cur + 1) % 3
                prog.add(op=OP_POOL2, dtype=self.dt, H=H, W=W, cin=ch, src=src, dst=bufs[nxt],
                         dst_f32=f32 if pools == n_pool else None)
                H, W = H // 2, W // 2
                cur, src = nxt, bufs[nxt]
        lat = self.buf(n, H * W, ch)
        prog.add(op=OP_SCALE, dtype=self.dt, H=H, W=W, cin=ch, src=f32, dst=lat)
        prog.add(op=OP_NHWC_OUT, dtype=self.dt, H=H, W=W, cin=ch, src=lat, dst=out_nchw)
        return prog

    # ------------------------------------------------------------------ MuZeroAgent-style calls (NCHW float32 tensors)
    def _latent_in(self, prog, h):
        H, W = self.latent_hw
        x = self.buf(h.shape[0], H * W, self.latent_ch)
        prog.add(op=OP_NCHW_IN, dtype=self.dt, H=H, W=W, cin=self.latent_ch, src=h, dst=x)
        return x

    def representation(self, state: torch.Tensor) -> torch.Tensor:
        x = state.to(self.device, torch.float32).contiguous()
        n = x.shape[0]
        out = torch.empty((n, self.latent_ch, *self.latent_hw), dtype=torch.float32, device=self.device)
        self.representation_program(n, x, out).run()
        return out

    create_hidden_state_root = representation

    def prediction(self, hidden: torch.Tensor):
        """-> (policy_logits (n,3), value_logits (n,11)) raw logits, like MuZeroAgent.evaluate_state."""
        h = hidden.to(self.device, torch.float32).contiguous()
        n = h.shape[0]
        H, W = self.latent_hw
        prog = Program(n)
        x = self._latent_in(prog, h)
        bufs = [self.buf(n, H * W, self.latent_ch) for _ in range(3)]
        mid = self.buf(n, H * W, self.latent_ch)
        pol = torch.empty((n, self.num_actions), dtype=torch.float32, device=self.device)
        val = torch.empty((n, self.num_supports), dtype=torch.float32, device=self.device)
        x32 = h.permute(0, 2, 3, 1).reshape(n, H * W, self.latent_ch).contiguous() if self.dt != F32 else None   # exact residual of the first block
        if x32 is not None:
            prog.keep.append(x32)
        prog.extend(self.prediction_program(n, x, bufs, mid, None, None, pol, val, value_mode=0, pi_mode=0, src_f32=x32))
        prog.run()
        return pol, val

    evaluate_state = prediction

    def dynamics(self, hidden: torch.Tensor, action_planes: torch.Tensor):
        """-> (scaled latent (n,256,4,5), reward_logits (n,11)), like MuZeroAgent.hidden_state_transition."""
        h = hidden.to(self.device, torch.float32).contiguous()
        n = h.shape[0]
        H, W = self.latent_hw
        act = action_planes.to(self.device)[:, :, 0, 0].argmax(dim=1).to(torch.int32).contiguous()   # one-hot planes (mcts.py:252-268)
        prog = Program(n)
        x = self._latent_in(prog, h)
        bufs = [self.buf(n, H * W, self.latent_ch) for _ in range(3)]
        mid = self.buf(n, H * W, self.latent_ch)
        f32 = self.buf(n, H * W, self.latent_ch, torch.float32)
        lat = self.buf(n, H * W, self.latent_ch)
        rew = torch.empty((n, self.num_supports), dtype=torch.float32, device=self.device)
        out = torch.empty((n, self.latent_ch, H, W), dtype=torch.float32, device=self.device)
        prog.extend(self.dynamics_program(n, x, act, bufs, mid, f32, None, lat, reward_logits=rew, reward_mode=0))
        prog.add(op=OP_NHWC_OUT, dtype=self.dt, H=H, W=W, cin=self.latent_ch, src=lat, dst=out)
        prog.run()
        return out, rew

    hidden_state_transition = dynamics

    def inverted_softmax_expectation(self, logits: torch.Tensor) -> torch.Tensor:
        """utils.py:74-81 (host-side torch ops; the search path uses the fused head kernel instead)."""
        s = torch.linspace(self.supports_min, self.supports_max, self.num_supports, device=logits.device)
        x = torch.sum(torch.softmax(logits, dim=-1) * s, dim=-1)
        return torch.sign(x) * ((torch.abs(x) + (1 - 0.001)) ** 2 - 1)
