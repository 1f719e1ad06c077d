"""The reference's training step on the device: loss and optimizer (csrc/train.cu: mz_loss / mz_adam), the
convolutions' data and weight gradients, training-mode BatchNorm, the ResidualBlock autograd bridge and the graphed loop iteration -- all
through the C ABI (include/mzb200.h).

    reference                                                      here
    loss_fn(observed_reward, predicted_reward, ...,                same call, same return tuple; ONE launch computes the three
            target_transformation, K)  train_torch.py:33-66        batch-mean KL divergences, the total and its gradient w.r.t. the
                                                                   three logit tensors; `loss.backward()` (train_torch.py:515) hands
                                                                   those stored gradients to autograd
    ScalarTransforms.supports_representation  utils.py:30-64       fused into the same launch (pass the bound method, the
                                                                   ScalarTransforms object or the supports tensor)
    MuZeroAgent.optimizer = torch.optim.Adam(self.parameters(),    Adam(module_or_params, lr, weight_decay): parameters and gradients
        lr, weight_decay=1e-4)  networks.py:268;                   re-pointed into two flat fp32 buffers, `step()` = ONE launch
        optimizer.zero_grad() / optimizer.step()  :497,516         (28 B per parameter, HBM-bound), `zero_grad()` = one memset

    conv backward w.r.t. its input (autograd of nn.Conv2d,         ConvDgrad(weight): the data gradient of a stride-1 "same" convolution IS a
        networks.py:11,24-25, inside loss.backward())              convolution of dY with the transposed, tap-flipped weights, so it runs on the
                                                                   acting path's tcgen05 kernel (csrc/conv_tc.cu) with re-packed weights

    weight gradients, training-mode BatchNorm, the autograd bridge    conv_wgrad / flush_wgrads (csrc/wgrad.cu), bn_train_forward / _backward
        of the ResidualBlock runs, the graphed loop iteration          (csrc/bn.cu), trunk_forward, GraphedTrainStep / accelerate_training_stage

The layers around the ResidualBlock runs (stems, pools, ConvBlocks, Linear heads, _scale_state) are in train_layers.py.  There is no CPU fallback.
"""
from __future__ import annotations

import os

import torch

from . import _lib


def _p(t):
    return None if t is None else t.data_ptr()


# 16-bit element type of the FORWARD operands of the training kernels (activations and weights of the convolutions).  fp16: 11 significant
# bits like TF32, which the reference's own training uses on a GPU (8 for bf16) -- the activations of these networks are O(1), far from
# fp16's range.  The gradients that feed the convolution gradients (dz) stay bf16: they need the exponent range.  MZB_TRAIN_FWD=bf16 reverts.
FWD_DTYPE = torch.bfloat16 if os.environ.get("MZB_TRAIN_FWD", "f16") == "bf16" else torch.float16


def _supports_of(target_transformation) -> torch.Tensor:
    if isinstance(target_transformation, torch.Tensor):
        return target_transformation
    owner = getattr(target_transformation, "__self__", target_transformation)   # bound supports_representation -> ScalarTransforms
    sup = getattr(owner, "supports", None)
    if sup is None:
        raise TypeError("target_transformation must be ScalarTransforms.supports_representation, a ScalarTransforms or the supports tensor")
    return sup


class _Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_reward, pred_value, pred_policy, obs_reward, value_target, visits, supports, K):
        _lib.require_cuda()
        dev = pred_reward.device
        if dev.type != "cuda":
            raise RuntimeError("loss_fn needs CUDA tensors (there is no CPU fallback)")
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        pr, pv, pp = f(pred_reward), f(pred_value), f(pred_policy)
        n_sup, n_act = pr.shape[-1], pp.shape[-1]
        rows = pr.numel() // n_sup
        o, v, c, s = f(obs_reward), f(value_target), f(visits), f(supports)
        if not (pv.shape == pr.shape and pp.numel() == rows * n_act and o.numel() == rows and v.numel() == rows
                and c.numel() == rows * n_act and s.numel() == n_sup):
            raise ValueError("loss_fn: inconsistent shapes")
        L = _lib.lib()
        scratch = torch.zeros((L.mz_loss_scratch_bytes(rows) + 7) // 8, dtype=torch.int64, device=dev)
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        d_r, d_v, d_p = torch.empty_like(pr), torch.empty_like(pv), torch.empty_like(pp)
        with torch.cuda.device(dev):
            _lib.check(L.mz_loss(rows, int(K), n_sup, n_act, _p(s), _p(pr), _p(pv), _p(pp), _p(o), _p(v), _p(c), _p(losses),
                                 _p(d_r), _p(d_v), _p(d_p), _p(scratch), torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(d_r, d_v, d_p)
        parts = tuple(losses[i] for i in (1, 2, 3))
        ctx.mark_non_differentiable(*parts)
        return (losses[0],) + parts

    @staticmethod
    def backward(ctx, g, *_):
        d_r, d_v, d_p = ctx.saved_tensors
        return g * d_r, g * d_v, g * d_p, None, None, None, None, None


def loss_fn(observed_reward, predicted_reward, bootstrapped_reward, predicted_value, visit_counts, predicted_policy,
            target_transformation, K):
    """Drop-in for train_torch.py:33-66.  Returns (loss, reward_loss, value_loss, policy_loss) as 0-dim CUDA tensors;
    `loss` carries the autograd edge to the three prediction tensors."""
    return _Loss.apply(predicted_reward, predicted_value, predicted_policy, observed_reward, bootstrapped_reward, visit_counts,
                       _supports_of(target_transformation), K)


class Adam:
    """torch.optim.Adam as networks.py:268 builds it (betas (0.9, 0.999), eps 1e-8, L2 weight decay added to the gradient), over
    flat buffers: every parameter's `.data` and `.grad` become views into `self.flat_param` / `self.flat_grad`, so autograd
    accumulates straight into the flat gradient and `step()` is one mz_adam launch."""

    def __init__(self, params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
        _lib.require_cuda()
        params = list(params.parameters()) if isinstance(params, torch.nn.Module) else list(params)
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("Adam: no parameters")
        dev = params[0].device
        if any(p.device != dev or p.dtype != torch.float32 for p in params) or dev.type != "cuda":
            raise RuntimeError("Adam: parameters must be float32 CUDA tensors on one device (there is no CPU fallback)")
        self.params, self.lr, self.betas, self.eps, self.weight_decay = params, float(lr), tuple(betas), float(eps), float(weight_decay)
        self.step_count = 0
        offs, n = [], 0
        for p in params:                      # 16-byte aligned segments; the padding holds zeros and stays zero
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros_like(self.flat_param)
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        with torch.no_grad():
            for p, o in zip(params, offs):
                seg = self.flat_param[o:o + p.numel()].view(p.shape)
                seg.copy_(p)
                gseg = self.flat_grad[o:o + p.numel()].view(p.shape)
                if p.grad is not None:
                    gseg.copy_(p.grad)
                p.data = seg
                p.grad = gseg
        self._offsets = offs
        self._skip_steps = 0
        self._dev_state = torch.zeros(16, dtype=torch.int32, device=dev)      # MZ_ADAM_STATE_BYTES: step count + this update's constants
        self._dev_step = 0

    def zero_grad(self, set_to_none: bool = False):
        """Zeros the flat gradient (the views stay attached: `set_to_none` is accepted and ignored)."""
        self.flat_grad.zero_()

    def step(self):
        if self._skip_steps > 0:                 # a graphed training step (GraphedTrainStep) has already applied this update
            self._skip_steps -= 1
            return
        self.step_count += 1
        dev = self.flat_param.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mz_adam(self.flat_param.numel(), _p(self.flat_param), _p(self.flat_grad), _p(self.exp_avg),
                                          _p(self.exp_avg_sq), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                          self.step_count, torch.cuda.current_stream(dev).cuda_stream))
        # the kernel writes through raw pointers: bump the parameters' version counters like an in-place torch op would, so that
        # anything keyed on them (MCTSSearchVec's re-pack check, autograd's saved-tensor checks) sees the update
        torch.autograd.graph.increment_version(self.params)

    def step_dev(self):
        """The same update with the step count read from (and advanced in) device memory -- mz_adam_dev -- so that it can sit inside a CUDA
        graph.  The caller keeps `step_count` in step with the replays (GraphedTrainStep does)."""
        dev = self.flat_param.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mz_adam_dev(self.flat_param.numel(), _p(self.flat_param), _p(self.flat_grad), _p(self.exp_avg), _p(self.exp_avg_sq),
                                              self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, _p(self._dev_state),
                                              torch.cuda.current_stream(dev).cuda_stream))

    def sync_dev_step(self):
        """device step counter := step_count (after eager steps or load_state_dict)"""
        if self._dev_step != self.step_count:
            self._dev_state[0] = self.step_count
            self._dev_step = self.step_count

    def state_dict(self):
        """torch.optim.Adam's own layout ({"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]}), so the checkpoint the
        reference writes (train_torch.py:624 `optimizer.state_dict()`) and reads (:652) is interchangeable with torch's optimizer."""
        state = {}
        if self.step_count > 0:
            for i, (p, o) in enumerate(zip(self.params, self._offsets)):
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts torch.optim.Adam's state_dict (one param group; every parameter's step must agree: the kernel keeps one step counter)
        and this class's earlier flat layout ("step", "exp_avg", "exp_avg_sq", ...)."""
        if "param_groups" not in sd:                       # flat layout of earlier versions
            self.step_count = int(sd["step"])
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            self.lr, self.betas, self.eps, self.weight_decay = float(sd["lr"]), tuple(sd["betas"]), float(sd["eps"]), float(sd["weight_decay"])
            return
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("Adam.load_state_dict: expected one parameter group over the same parameters")
        g = groups[0]
        if g.get("amsgrad") or g.get("maximize"):
            raise ValueError("Adam.load_state_dict: amsgrad / maximize are not built")
        self.lr, self.betas, self.eps, self.weight_decay = float(g["lr"]), tuple(g["betas"]), float(g["eps"]), float(g["weight_decay"])
        self.exp_avg.zero_(); self.exp_avg_sq.zero_()
        steps = set()
        for idx, (p, o) in zip(g["params"], zip(self.params, self._offsets)):
            st = sd["state"].get(idx)
            if st is None:
                steps.add(0)
                continue
            steps.add(int(float(st["step"])))
            n = p.numel()
            self.exp_avg[o:o + n].view(p.shape).copy_(st["exp_avg"])
            self.exp_avg_sq[o:o + n].view(p.shape).copy_(st["exp_avg_sq"])
        if len(steps) > 1:
            raise ValueError(f"Adam.load_state_dict: parameters have different step counts {sorted(steps)}; the flat update keeps one")
        self.step_count = steps.pop() if steps else 0


def pack_conv(weight: torch.Tensor, cin: int | None = None, fwd_dtype=None, fwd_out: torch.Tensor | None = None, dgrad_out: torch.Tensor | None = None):
    """The 16-bit operand packs of a convolution weight (cout, cin_total, k, k) -- its first `cin` input channels -- in ONE launch
    (mz_pack_conv): the forward pack [tap][cin/64][cout][64] in fwd_dtype (written to fwd_out, or allocated when fwd_dtype is given) and / or
    the data-gradient pack [tap][cout/64][cin][64] in bf16 (dgrad_out).  Returns (fwd pack or None, dgrad pack or None)."""
    _lib.require_cuda()
    cout, cin_total, k, _ = weight.shape
    cin = cin or cin_total
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    if fwd_out is None and fwd_dtype is not None:
        fwd_out = torch.empty((k * k, cin // 64, cout, 64), dtype=fwd_dtype, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(_lib.lib().mz_pack_conv(cout, cin_total, cin, k, _p(w), _p(fwd_out), _dt(fwd_out.dtype) if fwd_out is not None else 0, _p(dgrad_out),
                                           torch.cuda.current_stream(w.device).cuda_stream))
    return fwd_out, dgrad_out


class ConvDgrad:
    """dL/dx of `y = conv2d(x, weight, padding=k//2)` (stride 1; the only convolution form in networks.py) on the tensor cores:
    dx = conv2d(dy, weight.transpose(0, 1).flip(2, 3), padding=k//2), evaluated by the acting path's tcgen05 implicit-GEMM kernel with
    bf16 operands and fp32 accumulation.  Channels-last tensors: dy (n, H, W, cout) bf16 in, dx (n, H, W, cin) fp32 out."""

    def __init__(self, weight: torch.Tensor, device="cuda", cin: int | None = None):
        """cin: use only the first cin input channels of the weight (the dynamics ConvBlock's 256 hidden-state channels of 259)"""
        _lib.require_cuda()
        cout, cin_total, k, _ = weight.shape
        cin = cin or cin_total
        if k not in (1, 3) or cout % 64 or cin not in (128, 256):
            raise ValueError("ConvDgrad: built for the 1x1 / 3x3 convolutions of networks.py with cout % 64 == 0 and cin in (128, 256) (one UMMA N)")
        self.cin, self.cout, self.k = cin, cout, k
        # tile-contiguous [tap][cout/64][cin][64] of the transposed, tap-flipped filter, as networks.py _conv packs a forward weight
        self.w = torch.empty((k * k, cout // 64, cin, 64), dtype=torch.bfloat16, device=device)
        self.repack(weight)
        self.scale = torch.ones(cin, dtype=torch.float32, device=device)
        self.shift = torch.zeros(cin, dtype=torch.float32, device=device)

    def repack(self, weight: torch.Tensor):
        """re-pack in place from the (updated) weight: one mz_pack_conv launch"""
        pack_conv(weight.to(self.w.device), self.cin, None, None, self.w)

    @staticmethod
    def dgrad_filter(weight: torch.Tensor) -> torch.Tensor:
        """(cout, cin, k, k) forward filter -> (cin, cout, k, k) filter whose "same" convolution with dy is dx: channels swapped, taps flipped."""
        return weight.detach().to(torch.float32).transpose(0, 1).flip(2, 3)

    def __call__(self, dy: torch.Tensor, add: torch.Tensor | None = None) -> torch.Tensor:
        """add: float32 (n, H, W, cin) added in the kernel's epilogue (the skip connection's gradient of a ResidualBlock)"""
        from .src.networks import ACT, BF16, OP_CONV, Program
        n, H, W, c = dy.shape
        if c != self.cout or dy.dtype != torch.bfloat16 or not dy.is_cuda or not dy.is_contiguous():
            raise ValueError("ConvDgrad: dy must be a contiguous CUDA bf16 tensor (n, H, W, cout)")
        dx = torch.empty((n, H, W, self.cin), dtype=torch.float32, device=dy.device)
        prog = Program(n)                                # fp32-only output (dst = NULL): staged, coalesced stores (csrc/conv_tc.cu)
        extra = {}
        if add is not None:
            if add.shape != dx.shape or add.dtype != torch.float32 or not add.is_contiguous():
                raise ValueError("ConvDgrad: add must be a contiguous float32 tensor of the output's shape")
            extra["res_f32"] = add
        prog.add(op=OP_CONV, dtype=BF16, H=H, W=W, cin=self.cout, cout=self.cin, ksize=self.k, act=ACT["none"], use_tc=1, w_layout=1,
                 src=dy, dst_f32=dx, w=self.w, scale=self.scale, shift=self.shift, **extra)
        prog.run()
        return dx


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, ksize: int, accumulate_into: torch.Tensor | None = None) -> torch.Tensor:
    """dL/dweight of `y = conv2d(x, weight, padding=ksize//2)`, (cout, cin, k, k) float32, on the tensor cores (mz_conv_wgrad_any,
    csrc/wgrad.cu): cout in {128, 256}, cin in {64, 128, 256} -- the trunks' 256 -> 256 convolutions, the representation network's stems and
    128-channel blocks, the head ConvBlocks.  Channels-last 16-bit tensors: dy (n, H, W, cout), x (n, H, W, cin).
    accumulate_into: add the gradient to this tensor (a parameter's .grad) instead of returning a new one; it may have MORE input channels
    than x (cout, cin + extra, k, k): the gradient goes to the first cin (the dynamics ConvBlock's 256 hidden-state channels of 259)."""
    _lib.require_cuda()
    n, H, W, cin = x.shape
    cout = dy.shape[-1]
    # (dy bf16, x fp16) = a training step with fp16 forward operands: x is converted to bf16 inside its transpose
    combos = {(torch.bfloat16, torch.bfloat16): 1, (torch.float16, torch.float16): 2, (torch.bfloat16, torch.float16): 1}
    if (dy.shape[:3] != x.shape[:3] or cout not in (128, 256) or cin not in (64, 128, 256) or (dy.dtype, x.dtype) not in combos
            or not (x.is_cuda and dy.is_cuda)):
        raise ValueError("conv_wgrad: dy (n, H, W, cout in {128, 256}) and x (n, H, W, cin in {64, 128, 256}) must be CUDA tensors, both bf16, "
                         "both fp16, or dy bf16 with x fp16")
    L, dev = _lib.lib(), x.device
    ns = L.mz_wgrad_padded_samples(n)
    st = torch.cuda.current_stream(dev).cuda_stream
    dy_t = torch.empty((cout, H * W, ns), dtype=dy.dtype, device=dev)
    x_t = torch.empty((cin, H * W, ns), dtype=dy.dtype, device=dev)
    partial = torch.empty(L.mz_wgrad_partial_bytes_any(ksize, n, cout, cin) // 4, dtype=torch.float32, device=dev)
    acc = accumulate_into is not None
    if acc and not (accumulate_into.dim() == 4 and accumulate_into.shape[0] == cout and accumulate_into.shape[1] >= cin
                    and tuple(accumulate_into.shape[2:]) == (ksize, ksize) and accumulate_into.dtype == torch.float32
                    and accumulate_into.is_contiguous() and accumulate_into.device == dev):
        raise ValueError("conv_wgrad: accumulate_into must be a contiguous float32 (cout, >= cin, k, k) tensor on the operands' device")
    dw = accumulate_into if acc else torch.empty((cout, cin, ksize, ksize), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.mz_wgrad_transpose(n, H * W, cout, _p(dy.contiguous()), _p(dy_t), st))
        _lib.check(L.mz_wgrad_transpose_cvt(n, H * W, cin, _p(x.contiguous()), _p(x_t), int(x.dtype != dy.dtype), st))
        _lib.check(L.mz_conv_wgrad_any(n, H, W, ksize, combos[(dy.dtype, x.dtype)], cout, cin, dw.shape[1], _p(dy_t), _p(x_t), _p(partial), _p(dw),
                                       int(acc), st))
    return dw


# Deferred weight gradients.  The K unroll steps of a training step share every convolution (train_torch.py:507-525), so a weight's gradient is
# the sum of K small GEMMs (512 samples: 72 tiles of 100 k-steps, 23 us each at 340 TFLOP/s + a split reduction each).  Inside a graphed
# training step (GraphedTrainStep switches this on) the (dz, x) operand pairs are only queued during loss.backward(); flush_wgrads() then
# concatenates the pairs of a weight along the GEMM's reduction axis and runs ONE GEMM per weight (2560 samples: ~900 TFLOP/s, one split
# reduction, the K partial sums added in the fp32 accumulator instead of through K roundings of .grad).  Eager steps keep the immediate form:
# their .grad must be complete when backward() returns.
_DEFER = {"on": False, "q": {}, "expect": {}, "sides": []}
_COMBOS = {(torch.bfloat16, torch.bfloat16): 1, (torch.float16, torch.float16): 2, (torch.bfloat16, torch.float16): 1}


class _Pending:
    """the queued uses of one weight: `count` x (dy, x) pairs of `n` samples.  When the previous step has shown how many uses to expect, the
    K-concatenated operands exist from the first use on and every pair is transposed into its sample columns right away (on the weight-gradient
    side stream, next to the data-gradient chain); flush_wgrads() is then left with the GEMM."""

    def __init__(self, into, ksize, dy, x, expect):
        self.into, self.ksize = into, ksize
        self.shape = (dy.shape[0], dy.shape[1], dy.shape[2], dy.shape[3], x.shape[3], dy.dtype, x.dtype)
        self.items, self.keep, self.filled, self.bufs, self.count, self.done = [], [], 0, None, 0, False
        if expect is not None and expect[1] == self.shape and expect[0] > 0:
            n, H, W, cout, cin = self.shape[:5]
            self.ns = _lib.lib().mz_wgrad_padded_samples(n)
            self.count = expect[0]
            self.bufs = (torch.empty((cout, H * W, self.count * self.ns), dtype=dy.dtype, device=dy.device),
                         torch.empty((cin, H * W, self.count * self.ns), dtype=dy.dtype, device=dy.device))


def _transpose_pair(L, dy, x, dy_t, x_t, tot, off, st):
    n, H, W, cout = dy.shape
    if os.environ.get("MZB_TRAIN_TRANSPOSE_PAIR", "0") == "1":       # both operands in one launch (measured neutral: 30.3 vs 30.2 ms; off)
        _lib.check(L.mz_wgrad_transpose_pair(n, H * W, cout, _p(dy), _p(dy_t), 0, x.shape[3], _p(x), _p(x_t), int(x.dtype != dy.dtype), tot, off, st))
        return
    _lib.check(L.mz_wgrad_transpose_into(n, H * W, cout, _p(dy), _p(dy_t), tot, off, 0, st))
    _lib.check(L.mz_wgrad_transpose_into(n, H * W, x.shape[3], _p(x), _p(x_t), tot, off, int(x.dtype != dy.dtype), st))


def _pending_gemm(L, ent, dy_t, x_t, tot, stream, cur):
    n, H, W, cout, cin, dt_dy, dt_x = ent.shape
    dev = ent.into.device
    partial = torch.empty(L.mz_wgrad_partial_bytes_any(ent.ksize, tot, cout, cin) // 4, dtype=torch.float32, device=dev)
    _lib.check(L.mz_conv_wgrad_any(tot, H, W, ent.ksize, _COMBOS[(dt_dy, dt_x)], cout, cin, ent.into.shape[1], _p(dy_t), _p(x_t), _p(partial),
                                   _p(ent.into), 1, stream.cuda_stream))
    if stream is not cur:
        for t in (dy_t, x_t, partial):
            t.record_stream(stream)
    ent.keep.append(partial)


def wgrad_or_defer(dy16, x16, ksize, into):
    """conv_wgrad(dy16, x16, ksize, into) now, or -- in a graphed step, with an in-place destination -- queued for flush_wgrads()"""
    if not (_DEFER["on"] and into is not None):
        return conv_wgrad(dy16, x16, ksize, into)
    dy16, x16 = dy16.contiguous(), x16.contiguous()
    key = into.data_ptr()                                 # one entry per parameter (its .grad is a fixed view of the flat gradient buffer)
    ent = _DEFER["q"].get(key)
    if ent is None:
        ent = _DEFER["q"][key] = _Pending(into, ksize, dy16, x16, _DEFER["expect"].get(key))
    shape = (dy16.shape[0], dy16.shape[1], dy16.shape[2], dy16.shape[3], x16.shape[3], dy16.dtype, x16.dtype)
    if ent.bufs is not None and ent.filled < ent.count and shape == ent.shape:
        dev = dy16.device
        cur = torch.cuda.current_stream(dev)
        side = _wgrad_side_stream(cur) if _WGRAD_SIDE["on"] else None
        stream = side if side is not None else cur
        if side is not None:
            side.wait_stream(cur)
            if side not in _DEFER["sides"]:
                _DEFER["sides"].append(side)
        with torch.cuda.device(dev), torch.cuda.stream(stream):
            _transpose_pair(_lib.lib(), dy16, x16, ent.bufs[0], ent.bufs[1], ent.count * ent.ns, ent.filled * ent.ns, stream.cuda_stream)
            if ent.filled + 1 == ent.count and _DEFER.get("eager_gemm", False):
                # the last expected use: the GEMM follows its transposes on the same stream, next to what is left of the backward pass
                _pending_gemm(_lib.lib(), ent, ent.bufs[0], ent.bufs[1], ent.count * ent.ns, stream, cur)
                ent.done = True
        if side is not None:
            dy16.record_stream(side); x16.record_stream(side)
            for b in ent.bufs:
                b.record_stream(side)
        ent.keep.append((dy16, x16))                      # alive until the flush: their memory is not handed out again under the side stream's reads
        ent.filled += 1
    else:
        ent.items.append((dy16, x16))
    return None


def flush_wgrads(side=None):
    """Run the queued weight gradients: per weight ONE mz_conv_wgrad_any over the K-concatenated operand pair ([C][H*W][sum of padded sample
    counts]) that adds to the weight's .grad.  Pairs that were not transposed on arrival (first step: the number of uses was unknown) are
    transposed here.  side: a second stream; the weights alternate between it and the current stream."""
    q, _DEFER["q"] = _DEFER["q"], {}
    sides, _DEFER["sides"] = _DEFER["sides"], []
    if not q:
        return
    L = _lib.lib()
    dev = next(iter(q.values())).into.device
    cur = torch.cuda.current_stream(dev)
    for s_ in sides:                                      # the transposes issued during the backward pass
        cur.wait_stream(s_)
    if side is not None:
        side.wait_stream(cur)
    for i, ent in enumerate(q.values()):
        stream = side if (side is not None and i % 2 == 1) else cur
        n, H, W, cout, cin, dt_dy, dt_x = ent.shape
        with torch.cuda.device(dev), torch.cuda.stream(stream):
            st = stream.cuda_stream

            def gemm(dy_t, x_t, tot):
                _pending_gemm(L, ent, dy_t, x_t, tot, stream, cur)

            if ent.bufs is not None and not ent.done:
                tot = ent.count * ent.ns
                if ent.filled < ent.count:                # fewer uses than the step before: the unused sample columns must be zero
                    for b in ent.bufs:
                        b[:, :, ent.filled * ent.ns:].zero_()
                gemm(ent.bufs[0], ent.bufs[1], tot)
            if ent.items:
                ns = [L.mz_wgrad_padded_samples(dy.shape[0]) for dy, _ in ent.items]
                tot = sum(ns)
                dy_t = torch.empty((cout, H * W, tot), dtype=dt_dy, device=dev)
                x_t = torch.empty((cin, H * W, tot), dtype=dt_dy, device=dev)
                off = 0
                for (dy, x), n_i in zip(ent.items, ns):
                    if dy.shape[1:] != (H, W, cout) or x.shape[3] != cin or dy.dtype != dt_dy or x.dtype != dt_x:
                        raise ValueError("flush_wgrads: the operand pairs queued for one weight differ in shape or element type")
                    _transpose_pair(L, dy, x, dy_t, x_t, tot, off, st)
                    if stream is not cur:
                        dy.record_stream(stream); x.record_stream(stream)
                    off += n_i
                gemm(dy_t, x_t, tot)
        _DEFER["expect"][ent.into.data_ptr()] = (ent.filled + len(ent.items), ent.shape)
    if side is not None:
        cur.wait_stream(side)


_ACT = {"none": 0, "relu": 1, "leaky_relu": 2}


def _dt(t):
    return 2 if t == torch.float16 else 1


def rows16(x: torch.Tensor, dtype=None) -> torch.Tensor:
    """An NCHW-shaped float32 CUDA activation -> the kernels' contiguous channels-last 16-bit rows (n, H, W, C), on library kernels:
    mz_cvt16 when the tensor already has channels_last strides (what every bridge of this module returns), MZ_OP_NCHW_IN for a contiguous
    NCHW tensor (the representation network's input)."""
    _lib.require_cuda()
    from .src.networks import BF16, F16, OP_NCHW_IN, Program
    dtype = dtype or FWD_DTYPE
    x = x.detach()
    n, C_, H, W = x.shape
    out = torch.empty((n, H, W, C_), dtype=dtype, device=x.device)
    cl = x.permute(0, 2, 3, 1)
    if cl.is_contiguous() and x.dtype == torch.float32 and x.numel() % 4 == 0:
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mz_cvt16(x.numel(), _p(cl), _p(out), _dt(dtype), torch.cuda.current_stream(x.device).cuda_stream))
        return out
    x = x.float().contiguous()
    with torch.cuda.device(x.device):
        prog = Program(n)
        prog.add(op=OP_NCHW_IN, dtype=F16 if dtype == torch.float16 else BF16, H=H, W=W, cin=C_, src=x, dst=out)
        prog.run()
    return out


# BatchNorm statistics from the producing convolution's epilogue (mz_op.bn_partial) instead of a reduction pass over its output; MZB_TRAIN_CONV_STATS=0: off
CONV_STATS = os.environ.get("MZB_TRAIN_CONV_STATS", "1") == "1"


def conv_stats_buffer(n, H, W, ksize, cout, device):
    """the partial-sum blocks a training-form convolution writes for the BatchNorm that follows: float64 [blocks][2][cout]"""
    nb = _lib.lib().mz_conv_stats_blocks(n, H, W, ksize)
    return torch.empty((nb, 2, cout), dtype=torch.float64, device=device)


def bn_train_forward(z, gamma, beta, res=None, act="relu", eps=1e-5, momentum=0.1, running_mean=None, running_var=None, out_dtype=torch.bfloat16,
                     want16=True, stats=None):
    """Training-mode BatchNorm2d (+ residual) + activation of a ConvBlock / ResidualBlock (networks.py:16-17,31-35) on channels-last rows.
    z: float32 (..., C) convolution output incl. bias.  Returns (y 16-bit, y float32, save_mean, save_invstd); running stats updated in place."""
    _lib.require_cuda()
    L, dev, C_ = _lib.lib(), z.device, z.shape[-1]
    M = z.numel() // C_
    z = z.contiguous()
    y, y32 = (torch.empty(z.shape, dtype=out_dtype, device=dev) if want16 else None), torch.empty_like(z)
    mean, invstd = torch.empty(C_, device=dev), torch.empty(C_, device=dev)
    if stats is not None:                                  # the producing convolution has already written the partial sums (conv_stats)
        with torch.cuda.device(dev):
            _lib.check(L.mz_bn_train_fwd_pre(M, C_, stats.shape[0], _p(stats), _p(z), _p(gamma), _p(beta), _p(res), _dt(out_dtype), _ACT[act], eps, momentum,
                                             _p(running_mean), _p(running_var), _p(mean), _p(invstd), _p(y), _p(y32), torch.cuda.current_stream(dev).cuda_stream))
        return y, y32, mean, invstd
    scratch = torch.empty(L.mz_bn_scratch_bytes(M, C_) // 8, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.mz_bn_train_fwd(M, C_, _p(z), _p(gamma), _p(beta), _p(res), _dt(out_dtype), _ACT[act], eps, momentum, _p(running_mean), _p(running_var),
                                     _p(mean), _p(invstd), _p(y), _p(y32), _p(scratch), torch.cuda.current_stream(dev).cuda_stream))
    return y, y32, mean, invstd


def bn_train_backward(z, dy, gamma, beta, mean, invstd, res=None, act="relu", out_dtype=torch.bfloat16, want32=True, want_res=True, acc=(None, None)):
    """Backward of bn_train_forward.  dy: float32 gradient of the block output.  Returns (dz float32, dz 16-bit, dgamma, dbeta, dres float32).
    res keeps the element type the forward pass gave it (fp16 or bf16); out_dtype is that of the 16-bit dz."""
    _lib.require_cuda()
    L, dev, C_ = _lib.lib(), z.device, z.shape[-1]
    M = z.numel() // C_
    z, dy = z.contiguous(), dy.contiguous()
    dz = torch.empty_like(z) if want32 else None
    dz16 = torch.empty(z.shape, dtype=out_dtype, device=dev)
    dres = torch.empty_like(z) if want_res else None
    dgamma, dbeta = torch.empty(C_, device=dev), torch.empty(C_, device=dev)
    scratch = torch.empty(L.mz_bn_scratch_bytes(M, C_) // 8, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        # acc = (gamma.grad, beta.grad): the kernel also adds the two gradients to them (contiguous float32 [C])
        _lib.check(L.mz_bn_train_bwd_acc(M, C_, _p(z), _p(dy), _p(gamma), _p(beta), _p(res), _dt(res.dtype if res is not None else out_dtype), _dt(out_dtype),
                                         _ACT[act], _p(mean), _p(invstd), _p(dgamma), _p(dbeta), _p(acc[0]), _p(acc[1]), _p(dz), _p(dz16), _p(dres), _p(scratch),
                                         torch.cuda.current_stream(dev).cuda_stream))
    return dz, dz16, dgamma, dbeta, dres


class ResidualBlockTrain:
    """One training step through a ResidualBlock (networks.py:19-35: relu(bn2(conv2(relu(bn1(conv1 x)))) + x), 128 or 256 channels, train mode) made of
    this library's kernels only: tcgen05 convolutions (forward: conv_tc.cu; data gradient: the same kernel on the transposed, flipped weights;
    weight gradient: wgrad.cu) and the training-mode BatchNorm kernels (bn.cu).  Channels-last bf16 activations, fp32 statistics and gradients.
    The convolution biases get no gradient here: a BatchNorm follows each convolution and subtracts the batch mean, so d loss / d bias is zero
    (torch returns rounding noise for it)."""

    def __init__(self, conv1_w, conv1_b, bn1_w, bn1_b, conv2_w, conv2_b, bn2_w, bn2_b, device="cuda", eps=1e-5, momentum=0.1):
        _lib.require_cuda()
        f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        self.fwd_dtype = FWD_DTYPE
        self.C = C_ = int(conv1_w.shape[0])
        self.w = [conv1_w.detach(), conv2_w.detach()]
        self.wt = {self.fwd_dtype: [self._pack(w, device, self.fwd_dtype) for w in self.w]}     # forward: tile-contiguous [tap][cin/64][cout][64], 16-bit
        self.dgrad = [ConvDgrad(w.to(device), device) for w in self.w]
        self.b, self.gamma, self.beta = [f(conv1_b), f(conv2_b)], [f(bn1_w), f(bn2_w)], [f(bn1_b), f(bn2_b)]
        self.running_mean = [torch.zeros(C_, device=device) for _ in range(2)]
        self.running_var = [torch.ones(C_, device=device) for _ in range(2)]
        self.ones = torch.ones(C_, device=device)
        self.eps, self.momentum = eps, momentum
        self._saved = None

    @staticmethod
    def _pack(w, device, dtype=torch.bfloat16):
        cout, cin, k, _ = w.shape
        return w.detach().float().permute(0, 2, 3, 1).reshape(cout, k * k, cin // 64, 64).permute(1, 2, 0, 3).contiguous().to(device=device, dtype=dtype)

    def _weights(self, dtype):
        """the forward pack in the element type of the activations it meets (packed on first use per refresh)"""
        if dtype not in self.wt:
            self.wt[dtype] = [self._pack(w, self.ones.device, dtype) for w in self.w]
        return self.wt[dtype]

    def _conv(self, x16, i, with_stats=False):
        """z = conv_i(x16) + bias as float32; with_stats: (z, the BatchNorm partial sums its epilogue wrote, or None)"""
        from .src.networks import ACT, BF16, F16, OP_CONV, Program
        n, H, W, _ = x16.shape
        BF16 = F16 if x16.dtype == torch.float16 else BF16
        z = torch.empty(x16.shape, dtype=torch.float32, device=x16.device)
        stats = conv_stats_buffer(n, H, W, 3, self.C, x16.device) if (with_stats and CONV_STATS) else None
        prog = Program(n)
        prog.add(op=OP_CONV, dtype=BF16, H=H, W=W, cin=self.C, cout=self.C, ksize=3, act=ACT["none"], use_tc=1, w_layout=1, src=x16, dst_f32=z,
                 w=self._weights(x16.dtype)[i], scale=self.ones, shift=self.b[i], **({"bn_partial": stats} if stats is not None else {}))
        prog.run()
        return (z, stats) if with_stats else z

    def refresh(self, conv1_w, conv1_b, bn1_w, bn1_b, conv2_w, conv2_b, bn2_w, bn2_b, running=None):
        """Re-pack from the live parameters of a module when any of them changed (version counters: optimizer steps and load_state_dict bump
        them); `running` = ((mean1, var1), (mean2, var2)) aliases the module's BatchNorm buffers so the kernels update them in place."""
        params = (conv1_w, conv1_b, bn1_w, bn1_b, conv2_w, conv2_b, bn2_w, bn2_b)
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key != getattr(self, "_key", None):
            dev = self.ones.device
            f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
            self.w = [conv1_w.detach(), conv2_w.detach()]
            if all(w.is_cuda and w.device == dev for w in self.w):
                # both packs of a convolution in one launch, into the buffers of the step before (no allocation, no torch kernels)
                self.wt = {self.fwd_dtype: self.wt[self.fwd_dtype]}
                for i, w in enumerate(self.w):
                    pack_conv(w, None, None, self.wt[self.fwd_dtype][i], self.dgrad[i].w)
            else:
                self.wt = {self.fwd_dtype: [self._pack(w, dev, self.fwd_dtype) for w in self.w]}
                self.dgrad = [ConvDgrad(w.to(dev), dev) for w in self.w]
            self.b, self.gamma, self.beta = [f(conv1_b), f(conv2_b)], [f(bn1_w), f(bn2_w)], [f(bn1_b), f(bn2_b)]
            self._key = key
        if running is not None:
            self.running_mean, self.running_var = [running[0][0], running[1][0]], [running[0][1], running[1][1]]

    def forward_fn(self, x16: torch.Tensor):
        """Functional forward: returns (y bf16, y float32, saved) -- `saved` goes back into backward_fn, so one block object can be called
        several times per training step (the K unroll steps share their weights, train_torch.py:507-525)."""
        z1, st1 = self._conv(x16, 0, True)
        h16, _, m1, s1 = bn_train_forward(z1, self.gamma[0], self.beta[0], None, "relu", self.eps, self.momentum, self.running_mean[0], self.running_var[0],
                                          out_dtype=x16.dtype, stats=st1)
        z2, st2 = self._conv(h16, 1, True)
        y16, y32, m2, s2 = bn_train_forward(z2, self.gamma[1], self.beta[1], x16, "relu", self.eps, self.momentum, self.running_mean[1], self.running_var[1],
                                            out_dtype=x16.dtype, stats=st2)
        return y16, y32, (x16, z1, h16, z2, m1, s1, m2, s2)

    def forward(self, x16: torch.Tensor):
        """x16: (n, H, W, 256) bf16 channels-last.  Returns (y bf16, y float32)."""
        y16, y32, self._saved = self.forward_fn(x16)
        return y16, y32

    def backward(self, dy: torch.Tensor):
        """dy: float32 gradient of the block output.  Returns (dx float32, {parameter name: gradient})."""
        return self.backward_fn(dy, self._saved)

    def backward_fn(self, dy: torch.Tensor, saved, grad_into=(None, None), side=None, bn_into=((None, None), (None, None))):
        """grad_into: (conv1.weight.grad, conv2.weight.grad) to accumulate the weight gradients into (the entries of the returned dict are
        then None), or None entries for fresh tensors.  side: a CUDA stream for the weight-gradient chain (transposes, mz_conv_wgrad, split
        reduction) when both gradients accumulate in place -- it only shares its inputs with the data-gradient chain, and at a 512-sample
        minibatch neither fills the chip; the caller joins the stream."""
        x16, z1, h16, z2, m1, s1, m2, s2 = saved
        if side is not None and (grad_into[0] is None or grad_into[1] is None):
            side = None
        cur = torch.cuda.current_stream(dy.device) if side is not None else None

        def wgrad(dz16, act16, into):
            if _DEFER["on"] and into is not None:
                return wgrad_or_defer(dz16, act16, 3, into)
            if side is None:
                dw = conv_wgrad(dz16, act16, 3, into)
                return None if into is not None else dw
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                conv_wgrad(dz16, act16, 3, into)
            dz16.record_stream(side); act16.record_stream(side)
            return None

        # bn_into: ((bn1.weight.grad, bn1.bias.grad), (bn2...)) -- gradients the kernels add to in place (their dict entries are then None)
        _, dz2_16, dg2, db2, dres = bn_train_backward(z2, dy, self.gamma[1], self.beta[1], m2, s2, x16, "relu", want32=False, acc=bn_into[1])
        dw2 = wgrad(dz2_16, h16, grad_into[1])
        dh = self.dgrad[1](dz2_16)
        _, dz1_16, dg1, db1, _ = bn_train_backward(z1, dh, self.gamma[0], self.beta[0], m1, s1, None, "relu", want32=False, want_res=False, acc=bn_into[0])
        dw1 = wgrad(dz1_16, x16, grad_into[0])
        dx = self.dgrad[0](dz1_16, add=dres)            # + the skip connection's gradient, in the convolution's epilogue
        (g1a, b1a), (g2a, b2a) = bn_into
        return dx, {"conv1.weight": dw1, "bn1.weight": None if g1a is not None else dg1, "bn1.bias": None if b1a is not None else db1,
                    "conv2.weight": dw2, "bn2.weight": None if g2a is not None else dg2, "bn2.bias": None if b2a is not None else db2}


class TrunkTrain:
    """A run of train-mode ResidualBlocks (the 14-block trunks of the dynamics / prediction networks, networks.py:124-131,190-197) as one training
    step: forward keeps every block's saved tensors, backward walks the blocks in reverse.  `graph=True` captures forward + backward once as a
    CUDA graph (static shapes, fixed input / output-gradient buffers) and replays it."""

    def __init__(self, blocks):
        self.blocks = list(blocks)
        self._graph = None

    @classmethod
    def from_state_dict(cls, sd: dict, prefixes, device="cuda"):
        """prefixes: state_dict key prefixes of the ResidualBlocks, e.g. ["pred_net.res_blocks.0.", ...] (keys conv1/bn1/conv2/bn2 .weight/.bias)."""
        return cls(ResidualBlockTrain(*[sd[p + k] for k in ("conv1.weight", "conv1.bias", "bn1.weight", "bn1.bias", "conv2.weight", "conv2.bias",
                                                             "bn2.weight", "bn2.bias")], device=device) for p in prefixes)

    def forward(self, x16):
        y32 = None
        for b in self.blocks:
            x16, y32 = b.forward(x16)
        return x16, y32

    def backward(self, dy):
        grads = []
        for b in reversed(self.blocks):
            dy, g = b.backward(dy)
            grads.append(g)
        return dy, grads[::-1]

    def step(self, x16, dy, graph: bool = False):
        """forward(x16) then backward(dy).  Returns (y bf16, y float32, dx float32, [per-block gradient dicts])."""
        if not graph:
            y16, y32 = self.forward(x16)
            return (y16, y32) + self.backward(dy)
        if self._graph is None or self._in[0].shape != x16.shape:
            self._in = (torch.empty_like(x16), torch.empty_like(dy))
            self._in[0].copy_(x16); self._in[1].copy_(dy)
            side = torch.cuda.Stream(device=x16.device)
            side.wait_stream(torch.cuda.current_stream(x16.device))
            with torch.cuda.stream(side):                                  # warm-up outside the capture (first-use attribute calls, allocator)
                for _ in range(2):
                    self.forward(self._in[0]); self.backward(self._in[1])
            torch.cuda.current_stream(x16.device).wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                y16, y32 = self.forward(self._in[0])
                self._out = (y16, y32) + self.backward(self._in[1])
        self._in[0].copy_(x16); self._in[1].copy_(dy)
        self._graph.replay()
        return self._out


# ------------------------------------------------------------------------------------------------------------------------------------
# autograd bridge: a run of train-mode ResidualBlocks of an nn.Module (the reference's own ResidualBlock modules, networks.py:19-35, or the
# drop-in agent's) evaluated by this library's kernels inside loss.backward() (train_torch.py:515)
# The representation network's 256-channel blocks at 16x20 / 8x10 run on the same kernels (any image size: same errors as at 4x5,
# profiles/prof_train_any_hw.py); MZB_TRAIN_ANY_HW=0 keeps them on torch ops.  Fidelity against the fp32 modules on the whole K-step rollout
# (tests/test_train_agent_gpu.py prints it): forward outputs 0.022-0.026 of range and worst trunk weight-gradient cosine 0.93 with them,
# 0.010-0.013 and 0.98 without -- torch's TF32 convolutions, what the reference trains with on a GPU: 0.05-0.06 and 0.87; autocast bf16:
# 0.38-0.56 and 0.35.  (With bf16 forward operands, the build before FWD_DTYPE: 0.18-0.21 / 0.65 and 0.07-0.09 / 0.84.)
_ANY_HW = os.environ.get("MZB_TRAIN_ANY_HW", "1") == "1"


# GraphedTrainStep switches this on while it warms up / captures: inside a CUDA graph a stream fork costs nothing, eagerly the event
# record / wait pairs would eat the gain
_WGRAD_SIDE = {"on": False, "streams": {}}


def _wgrad_side_stream(cur):
    key = (cur.device, cur.cuda_stream)
    st = _WGRAD_SIDE["streams"].get(key)
    if st is None:
        st = _WGRAD_SIDE["streams"][key] = torch.cuda.Stream(device=cur.device)
    return st


class _TrunkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernels, *params):
        # channels-last 16-bit, the kernels' layout (a channels_last input -- what the drop-in agent's layers produce -- is permuted for free)
        x16 = rows16(x)
        saved, y32 = [], None
        for blk in kernels:
            x16, y32, sv = blk.forward_fn(x16)
            saved.append(sv)
        ctx.kernels, ctx.saved_blocks, ctx.params = kernels, saved, params
        return y32.permute(0, 3, 1, 2)                     # NCHW shape, channels_last strides: no copy

    @staticmethod
    def backward(ctx, dy):
        g = dy.permute(0, 2, 3, 1).contiguous().float()
        flat = []
        nb = len(ctx.kernels)
        cur = torch.cuda.current_stream(g.device)
        # (deferred weight gradients: nothing runs on the side stream here, and a stream that never joined the capture must not be waited for)
        side = _wgrad_side_stream(cur) if (_WGRAD_SIDE["on"] and not _DEFER["on"]) else None
        for i, (blk, sv) in enumerate(zip(reversed(ctx.kernels), reversed(ctx.saved_blocks))):
            ps = ctx.params[8 * (nb - 1 - i):8 * (nb - i)]           # conv1.w, conv1.b, bn1.w, bn1.b, conv2.w, conv2.b, bn2.w, bn2.b
            # a parameter that already has a contiguous .grad (always, under this module's flat-buffer Adam) takes its weight gradient by
            # in-kernel accumulation; autograd then gets None for it (one add kernel less per convolution and unroll step)
            slot = lambda p: p.grad if (p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32) else None
            into = (slot(ps[0]), slot(ps[4]))
            bn_into = ((slot(ps[2]), slot(ps[3])), (slot(ps[6]), slot(ps[7])))
            g, grads = blk.backward_fn(g, sv, into, side, bn_into)
            # conv biases: a BatchNorm follows and subtracts the batch mean (exactly zero gradient): None where a .grad exists, zeros otherwise
            zb = [None if p.grad is not None else torch.zeros(p.shape[0], device=g.device) for p in (ps[1], ps[5])]
            flat.append((grads["conv1.weight"], zb[0], grads["bn1.weight"], grads["bn1.bias"], grads["conv2.weight"], zb[1], grads["bn2.weight"], grads["bn2.bias"]))
        ctx.saved_blocks = None
        if side is not None:
            cur.wait_stream(side)                            # the accumulated weight gradients are complete when this node is
        out = [t for blk in reversed(flat) for t in blk]
        return (g.permute(0, 3, 1, 2), None, *out)


def _block_params(m):
    return (m.conv1.weight, m.conv1.bias, m.bn1.weight, m.bn1.bias, m.conv2.weight, m.conv2.bias, m.bn2.weight, m.bn2.bias)


def trunk_supported(blocks, x) -> bool:
    """Can this run of ResidualBlock modules go through the library kernels?  128 or 256 channels, 3x3 convolutions, ReLU, BatchNorm with
    running statistics at the default eps, a float32 CUDA input, training mode with gradients enabled."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] in (128, 256) and torch.is_grad_enabled()):
        return False
    C_ = x.shape[1]
    if x.shape[2] * x.shape[3] != 20 and not _ANY_HW:        # other maps than the 4x5 latent: the representation network's 16x20 / 8x10 blocks
        return False
    for m in blocks:
        if not (m.training and all(hasattr(m, a) for a in ("conv1", "bn1", "conv2", "bn2")) and m.conv1.weight.shape == (C_, C_, 3, 3)
                and m.conv2.weight.shape == (C_, C_, 3, 3) and isinstance(getattr(m, "act", None), torch.nn.ReLU)
                and m.bn1.track_running_stats and m.bn1.eps == 1e-5 and m.bn2.eps == 1e-5 and m.bn1.momentum == m.bn2.momentum and m.bn1.momentum is not None):
            return False
    return len(blocks) > 0


def trunk_forward(blocks, x: torch.Tensor) -> torch.Tensor:
    """`for b in blocks: x = b(x)` for train-mode 256-channel ResidualBlock modules, forward AND backward on this library's kernels (tcgen05
    convolution / data gradient / weight gradient, training-mode BatchNorm kernels): a differentiable torch op whose gradients flow to the
    modules' own Parameters; the modules' running statistics and num_batches_tracked are updated like nn.BatchNorm2d does."""
    kernels = []
    for m in blocks:
        k = getattr(m, "_mzb_kernels", None)
        if k is None:
            k = ResidualBlockTrain(*_block_params(m), device=x.device, eps=m.bn1.eps, momentum=m.bn1.momentum)
            object.__setattr__(m, "_mzb_kernels", k)          # not a submodule / buffer: invisible to state_dict()
        k.refresh(*_block_params(m), running=((m.bn1.running_mean, m.bn1.running_var), (m.bn2.running_mean, m.bn2.running_var)))
        kernels.append(k)
    with torch.no_grad():                                  # one multi-tensor kernel instead of two tiny ones per block
        torch._foreach_add_([b.num_batches_tracked for m in blocks for b in (m.bn1, m.bn2)], 1)
    params = [p for m in blocks for p in _block_params(m)]
    return _TrunkFn.apply(x, kernels, *params)


def accelerate_agent(agent):
    """Patch a MuZeroAgent-shaped module (the reference's own, networks.py:245-350) in place: in training mode on a CUDA device its three
    networks run forward AND backward on this library's kernels -- the ResidualBlock runs through trunk_forward, the stems, pools,
    ConvBlocks (incl. the dynamics ConvBlock's action planes), Linear heads and `_scale_state` through train_layers.py -- by binding the
    drop-in agent's forwards (src/agent.py; they only use the reference's attribute names) to the reference's modules; whatever a bridge
    does not take (eval mode, no_grad, other shapes) stays on the module's own torch ops.  The optimizer becomes this library's flat-buffer
    Adam with torch's hyper-parameters.  Returns the agent."""
    import types
    from .src import agent as A

    agent.rep_net.forward = types.MethodType(A.RepresentationNetwork.forward, agent.rep_net)
    agent.dyn_net.forward = types.MethodType(A.DynamicsNetwork.forward, agent.dyn_net)
    agent.pred_net.forward = types.MethodType(A.PredictionNetwork.forward, agent.pred_net)
    agent.hidden_state_transition = types.MethodType(A.MuZeroAgent.hidden_state_transition, agent)
    agent._scale_state = types.MethodType(A.MuZeroAgent._scale_state, agent)
    old = getattr(agent, "optimizer", None)
    if isinstance(old, torch.optim.Adam) and next(agent.parameters()).is_cuda:
        g = old.param_groups[0]
        agent.optimizer = Adam(agent.parameters(), lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"])
    return agent


# ------------------------------------------------------------------------------------------------------------------------------------
# one iteration of the training loop as ONE CUDA graph
def k_step_rollout(agent, input_states, input_actions, k_step_actions, K, latent_resolution=(4, 5), n_actions=3, side_stream=None):
    """What RLSystem._k_step_rollout does (train_torch.py:487-528) with any MuZeroAgent-shaped module: representation network on
    cat(states, action planes), then K x (prediction, dynamics on the one-hot action planes of :295-311).  Returns the stacked
    (reward, value, policy) logits, (batch, K, .).
    side_stream: evaluate the prediction network of step k on this stream while the dynamics network of step k runs on the current one
    (both only read h_k; autograd replays the split in the backward pass): at a 512-sample minibatch a trunk convolution fills 80 of the
    148 SMs, so the two networks overlap instead of queueing."""
    h = agent.create_hidden_state_root(torch.cat((input_states, input_actions), dim=1))
    cur = torch.cuda.current_stream(h.device) if h.is_cuda else None
    pol, val, rew = [], [], []
    for k in range(K):
        planes = torch.nn.functional.one_hot(k_step_actions[:, k].long(), num_classes=n_actions).float().view(-1, n_actions, 1, 1)
        if side_stream is not None:
            side_stream.wait_stream(cur)
            with torch.cuda.stream(side_stream):
                p_, v_ = agent.evaluate_state(h)
            h.record_stream(side_stream)
        else:
            p_, v_ = agent.evaluate_state(h)
        h, r_ = agent.hidden_state_transition(h, planes.expand(-1, -1, latent_resolution[0], latent_resolution[1]))
        pol.append(p_); val.append(v_); rew.append(r_)
    if side_stream is not None:
        cur.wait_stream(side_stream)
        for t in pol + val:
            t.record_stream(cur)
    return torch.stack(rew, dim=1), torch.stack(val, dim=1), torch.stack(pol, dim=1)


def _trunk_kernels(agent):
    return [m._mzb_kernels for m in agent.modules() if getattr(m, "_mzb_kernels", None) is not None]


class GraphedTrainStep:
    """The body of the reference's training loop (train_torch.py:385-417: zero_grad, _k_step_rollout, loss_fn, loss.backward(),
    optimizer.step()) captured once per minibatch shape as ONE CUDA graph and replayed: ~3 400 library launches + ~1 500 torch kernels
    of an eager step become one graph launch (the eager step is launch-bound: DESIGN.md section 9).  `agent` is the learner-side drop-in
    (src/agent.py) or a reference MuZeroAgent after accelerate_agent(); its optimizer must be this module's Adam (the step count lives on
    the device, mz_adam_dev).  Same arithmetic as the eager step: the captured kernels are the ones the eager path launches.

        step = GraphedTrainStep(agent, scalar_transforms.supports_representation, K)
        loss, reward_loss, value_loss, policy_loss = step(states, action_planes, k_actions, k_rewards, k_values, k_visit_counts)

    Capture needs one warm-up pass (kernel attributes, cuDNN plans, allocator); it runs on a snapshot: parameters, BatchNorm buffers and
    optimizer state are restored before the capture, so the first call counts as exactly one update."""

    def __init__(self, agent, target_transformation, K, latent_resolution=(4, 5), n_actions=3, rollout=None, two_streams=True):
        _lib.require_cuda()
        if not isinstance(getattr(agent, "optimizer", None), Adam):
            raise TypeError("GraphedTrainStep needs the flat-buffer Adam (the drop-in agent, or accelerate_agent(reference_agent))")
        self.agent, self.opt, self.K = agent, agent.optimizer, int(K)
        self.supports = _supports_of(target_transformation)
        self.res, self.n_actions = tuple(latent_resolution), int(n_actions)
        # MZB_TRAIN_STREAMS=1: everything on one stream
        self.side = torch.cuda.Stream(device=agent.optimizer.flat_param.device) if (two_streams and os.environ.get("MZB_TRAIN_STREAMS", "2") != "1") else None
        self.rollout = rollout or (lambda *inp: k_step_rollout(self.agent, *inp, self.K, self.res, self.n_actions, side_stream=self.side))
        self._graphs = {}
        self.replays = 0

    def _body(self, st):
        _WGRAD_SIDE["on"] = self.side is not None
        _DEFER["on"], _DEFER["q"], _DEFER["sides"] = os.environ.get("MZB_TRAIN_DEFER_WGRAD", "1") == "1", {}, []
        _DEFER["eager_gemm"] = os.environ.get("MZB_TRAIN_WGRAD_AT_LAST_USE", "0") == "1"      # measured: 36.3 vs 35.7 ms with the GEMMs at the end
        try:
            return self._body_inner(st)
        finally:
            _WGRAD_SIDE["on"] = False
            _DEFER["on"], _DEFER["q"], _DEFER["sides"] = False, {}, []

    def _body_inner(self, st):
        self.opt.flat_grad.zero_()
        pr, pv, pp = self.rollout(st[0], st[1], st[2])
        out = loss_fn(st[3], pr, st[4], pv, st[5], pp, self.supports, self.K)
        out[0].backward()
        flush_wgrads(self.side)                          # one weight-gradient GEMM per convolution over all its K uses
        self.opt.step_dev()
        return torch.stack([o.detach() for o in out])

    def _capture(self, inputs):
        dev = self.opt.flat_param.device
        st = [torch.empty_like(t, device=dev) for t in inputs]
        for d, t in zip(st, inputs):
            d.copy_(t)
        # warm-up on a snapshot of everything a step mutates
        sd = {k: v.clone() for k, v in self.agent.state_dict().items()}
        m, v, steps = self.opt.exp_avg.clone(), self.opt.exp_avg_sq.clone(), self.opt.step_count
        self.opt.sync_dev_step()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._body(st)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.agent.load_state_dict(sd)
        self.opt.exp_avg.copy_(m); self.opt.exp_avg_sq.copy_(v)
        self.opt.step_count = steps
        self.opt._dev_step = -1
        self.opt.sync_dev_step()
        for k in _trunk_kernels(self.agent):           # the weight re-packs must be part of the graph: they run before every replayed step
            k._key = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self._body(st)
        return g, st, out

    def __call__(self, input_states, input_actions, k_step_actions, k_step_rewards, k_step_values, k_step_policies):
        inputs = (input_states, input_actions, k_step_actions, k_step_rewards, k_step_values, k_step_policies)
        key = tuple((tuple(t.shape), t.dtype) for t in inputs)
        if key not in self._graphs:
            self._graphs[key] = self._capture(inputs)
        g, st, out = self._graphs[key]
        for d, t in zip(st, inputs):
            d.copy_(t, non_blocking=True)
        self.opt.sync_dev_step()
        g.replay()
        self.opt.step_count += 1
        self.opt._dev_step += 1
        self.replays += 1
        torch.autograd.graph.increment_version(self.opt.params)        # what the in-place update would have done (MCTSSearchVec's re-pack check)
        res = out.clone()
        return res[0], res[1], res[2], res[3]


def accelerate_training_stage(system, trainer_module=None):
    """Make the reference's UNMODIFIED `RLSystem._training_stage` (train_torch.py:369-452) run each loop iteration as one GraphedTrainStep
    replay.  The loop body is eager Python that calls, in order, optimizer.zero_grad(), self._k_step_rollout(...), the module-global
    loss_fn(...), loss.item(), loss.backward(), optimizer.step(); this patches three of those plug points on the live objects:
      * `system._k_step_rollout` only records its inputs and returns placeholders;
      * the trainer module's `loss_fn` -- when it is handed those placeholders -- replays the graph (rollout, loss, backward, Adam) with the
        recorded inputs and the targets it receives, and returns the four losses (`loss.backward()` on the result is then a no-op);
      * the following `optimizer.step()` is skipped once (the graph has applied the update).
    `system.mu_zero` must be the drop-in agent or an accelerate_agent()-ed reference agent.  Returns the GraphedTrainStep."""
    import sys
    mod = trainer_module or sys.modules[type(system).__module__]
    agent = system.mu_zero
    step = GraphedTrainStep(agent, system.scalar_transforms.supports_representation, system.K, tuple(system.latent_resolution), system.n_actions)
    marker = tuple(torch.empty(0) for _ in range(3))
    pending = {}
    eager_loss = mod.loss_fn

    def deferred_rollout(input_states, input_actions, k_step_actions):
        pending["inputs"] = (input_states, input_actions, k_step_actions)
        return marker

    def graphed_loss_fn(observed_reward, predicted_reward, bootstrapped_reward, predicted_value, visit_counts, predicted_policy, target_transformation, K):
        if predicted_reward is not marker[0]:
            return eager_loss(observed_reward=observed_reward, predicted_reward=predicted_reward, bootstrapped_reward=bootstrapped_reward,
                              predicted_value=predicted_value, visit_counts=visit_counts, predicted_policy=predicted_policy,
                              target_transformation=target_transformation, K=K)
        states, planes, acts = pending.pop("inputs")
        loss, rl, vl, pl = step(states, planes, acts, observed_reward, bootstrapped_reward, visit_counts)
        agent.optimizer._skip_steps = 1
        return loss.clone().requires_grad_(True), rl, vl, pl          # a leaf: the caller's loss.backward() touches nothing else

    system._k_step_rollout = deferred_rollout
    mod.loss_fn = graphed_loss_fn
    return step
