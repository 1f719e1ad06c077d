"""Build + ctypes loader of libmzb200.so (the C ABI declared in include/mzb200.h)."""
from __future__ import annotations

import ctypes as C
import glob
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
_SO = os.path.join(_PKG, "libmzb200.so")
_HDR = os.path.join(os.path.dirname(_PKG), "include", "mzb200.h")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

_lib = None


def _sources():
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    deps = _sources() + glob.glob(os.path.join(_CSRC, "*.cuh")) + [_HDR]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> muzero-breakout_b200/libmzb200.so (in-tree).  One object per source under
    csrc/_obj/ (git-ignored), compiled in parallel and re-used while neither the source nor a header changed; `force` recompiles all."""
    if not (force or _stale()):
        return _SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("libmzb200.so is missing/stale and nvcc was not found; there is no CPU fallback")
    from concurrent.futures import ThreadPoolExecutor

    objdir = os.path.join(_CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(d) for d in glob.glob(os.path.join(_CSRC, "*.cuh")) + [_HDR])
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, ""
        r = subprocess.run([nvcc] + cflags + ["-c", src, "-o", obj + ".tmp.o"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
        os.replace(obj + ".tmp.o", obj)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        done = list(ex.map(compile_one, _sources()))
    r = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", _SO + ".tmp"] + [o for o, _ in done] + ["-lcuda"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    os.replace(_SO + ".tmp", _SO)
    if verbose:
        print("".join(e for _, e in done))
    return _SO


class TreeArgs(C.Structure):
    """mirror of struct mz_tree_args (include/mzb200.h)"""
    _fields_ = [("B", C.c_int32), ("num_simulations", C.c_int32), ("sim", C.c_int32), ("reserved", C.c_int32),
                ("trees", C.c_void_p), ("s_tab", C.c_void_p), ("k_tab", C.c_void_p),
                ("discount", C.c_double), ("noise_weight", C.c_double), ("seed", C.c_uint64),
                ("reward", C.c_void_p), ("value", C.c_void_p), ("pi", C.c_void_p), ("noise", C.c_void_p),
                ("leaf_parent", C.c_void_p), ("leaf_action", C.c_void_p), ("leaf_slot", C.c_void_p),
                ("latent_store", C.c_void_p), ("dyn_in", C.c_void_p), ("latent_bytes", C.c_int64),
                ("out_value", C.c_void_p), ("out_visits", C.c_void_p), ("depth_hist", C.c_void_p),
                ("seed_dev", C.c_void_p)]


def _declare(L: C.CDLL) -> None:
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    L.mzb_version.restype = i32
    L.mzb_last_error.restype = C.c_char_p
    L.mzb_launch_count.restype = u64
    L.mzb_sizeof.argtypes, L.mzb_sizeof.restype = [i32], C.c_size_t
    sig = {
        "bk_env_reset": [i32] + [vp] * 8,
        "bk_env_reset_device_rng": [i32, vp, vp, u64, u64, vp, vp],
        "bk_env_step": [i32] + [vp] * 11,
        "bk_env_step_host": [i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp],
        "bk_env_ingest": [i32] + [vp] * 7,
        "bk_env_render": [i32] + [vp] * 4,
        "bk_env_velocity": [i32] + [vp] * 4,
        "bk_gray": [i32] + [vp] * 3,
    }
    L.bk_env_io_layout.argtypes, L.bk_env_io_layout.restype = [i32, i32, i32, vp], C.c_size_t
    L.mz_tree_bytes.argtypes, L.mz_tree_bytes.restype = [i32], C.c_size_t
    L.mz_tree_nodes.argtypes, L.mz_tree_nodes.restype = [i32], i32
    L.mz_stack_layer_bytes.argtypes, L.mz_stack_layer_bytes.restype = [], C.c_size_t
    L.mz_lat_layer_bytes.argtypes, L.mz_lat_layer_bytes.restype = [], C.c_size_t
    L.mz_stack_scratch_bytes.argtypes, L.mz_stack_scratch_bytes.restype = [i32, i32], C.c_size_t
    L.mz_conv_lo_bytes.argtypes, L.mz_conv_lo_bytes.restype = [i32, i32, i32, i32, i32], C.c_size_t
    L.mz_lat_max_samples.argtypes, L.mz_lat_max_samples.restype = [], i32
    L.mz_lat_max_layers.argtypes, L.mz_lat_max_layers.restype = [], i32
    L.mz_lat_scratch_bytes.argtypes, L.mz_lat_scratch_bytes.restype = [i32], C.c_size_t
    sig.update({
        "mz_puct_tables": [i32, C.c_double, C.c_double, vp, vp],
        "mz_tree_root": [C.POINTER(TreeArgs), vp],
        "mz_tree_step": [C.POINTER(TreeArgs), vp],
        "mz_run": [vp, i32, i32, vp],
        "mz_rep_input": [i32, i32, vp, i32, vp, vp, i32, vp, i32, vp],
        "mz_stack_build": [vp, i32, vp, C.c_size_t, vp, i32],
        "mz_stack_run": [vp, i32, i32, i32, i32, vp, i32, vp, vp, i32, vp],
        "mz_lat_build": [vp, i32, vp, C.c_size_t],
        "mz_lat_run": [vp, i32, i32, i32, vp, vp, i32, vp],
        "mz_sample_actions": [i32, vp, C.c_double, u64, C.c_uint32, vp, vp, vp, vp],
    })
    L.rb_plan_bytes.argtypes, L.rb_plan_bytes.restype = [i32], C.c_size_t
    L.rb_entries_for.argtypes, L.rb_entries_for.restype = [i32, i32, i32], C.c_longlong
    sig.update({
        "rb_append": [vp, i32, i32] + [vp] * 6 + [i32, vp, vp, i32, i32, vp, vp, vp],
        "rb_gather": [vp, i32] + [vp] * 11,
        "rb_gather_input": [vp, i32, vp, i32, vp, vp, vp],
    })
    L.mz_loss_scratch_bytes.argtypes, L.mz_loss_scratch_bytes.restype = [i32], C.c_size_t
    sig.update({
        "mz_loss": [i32, i32, i32, i32] + [vp] * 13,
        "mz_adam": [C.c_longlong, vp, vp, vp, vp] + [C.c_double] * 5 + [i32, vp],
        "mz_adam_dev": [C.c_longlong, vp, vp, vp, vp] + [C.c_double] * 5 + [vp, vp],
        "mz_wgrad_transpose": [i32, i32, i32, vp, vp, vp],
        "mz_wgrad_transpose_cvt": [i32, i32, i32, vp, vp, i32, vp],
        "mz_wgrad_transpose_into": [i32, i32, i32, vp, vp, i32, i32, i32, vp],
        "mz_wgrad_transpose_pair": [i32, i32, i32, vp, vp, i32, i32, vp, vp, i32, i32, i32, vp],
        "mz_conv_wgrad": [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp],
        "mz_conv_wgrad_accum": [i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp],
    })
    sig.update({
        "mz_bn_train_fwd": [i32, i32, vp, vp, vp, vp, i32, i32, C.c_double, C.c_double] + [vp] * 8,
        "mz_bn_train_bwd": [i32, i32, vp, vp, vp, vp, vp, i32, i32] + [vp] * 9,
        "mz_bn_train_bwd_mixed": [i32, i32, vp, vp, vp, vp, vp, i32, i32, i32] + [vp] * 9,
        "mz_bn_train_fwd_pre": [i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, C.c_double, C.c_double] + [vp] * 7,
        "mz_bn_train_bwd_acc": [i32, i32, vp, vp, vp, vp, vp, i32, i32, i32] + [vp] * 11,
    })
    ll = C.c_longlong
    sig.update({
        "mz_conv_wgrad_any": [i32] * 8 + [vp, vp, vp, vp, i32, vp],
        "mz_colsum": [i32, i32, vp, vp, i32, vp, vp],
        "mz_cvt16": [ll, vp, vp, i32, vp],
        "mz_pack_conv": [i32, i32, i32, i32, vp, vp, i32, vp, vp],
        "mz_pool2_train_fwd": [i32, i32, i32, i32, vp, vp, vp, i32, vp],
        "mz_pool2_train_bwd": [i32, i32, i32, i32, vp, vp, vp],
        "mz_linear_fwd": [i32, i32, i32, i32, vp, vp, vp, vp, vp],
        "mz_linear_bwd": [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp],
        "mz_scale_train_fwd": [i32, i32, vp, vp, vp, i32, vp, vp],
        "mz_scale_train_bwd": [i32, i32, vp, vp, vp, vp, vp],
        "mz_planes_conv_fwd": [i32] * 7 + [vp, ll, ll, ll, ll, vp, vp, vp],
        "mz_planes_conv_wgrad": [i32] * 7 + [vp, ll, ll, ll, ll, vp, vp, i32, vp, vp],
    })
    L.mz_wgrad_partial_bytes_any.argtypes, L.mz_wgrad_partial_bytes_any.restype = [i32, i32, i32, i32], C.c_size_t
    L.mz_linear_scratch_bytes.argtypes, L.mz_linear_scratch_bytes.restype = [i32, i32, i32, i32], C.c_size_t
    L.mz_planes_wgrad_scratch_bytes.argtypes, L.mz_planes_wgrad_scratch_bytes.restype = [i32, i32], C.c_size_t
    L.mz_bn_scratch_bytes.argtypes, L.mz_bn_scratch_bytes.restype = [i32, i32], C.c_size_t
    L.mz_conv_stats_blocks.argtypes, L.mz_conv_stats_blocks.restype = [i32, i32, i32, i32], i32
    L.mz_wgrad_padded_samples.argtypes, L.mz_wgrad_padded_samples.restype = [i32], i32
    L.mz_wgrad_partial_bytes.argtypes, L.mz_wgrad_partial_bytes.restype = [i32, i32], C.c_size_t
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes, f.restype = args, i32


def lib() -> C.CDLL:
    """The loaded library.  Raises (never falls back) when it cannot be built or loaded."""
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        _declare(L)
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"libmzb200 error {rc}: {lib().mzb_last_error().decode()}")


def launch_count() -> int:
    return int(lib().mzb_launch_count())


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("muzero_breakout_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
