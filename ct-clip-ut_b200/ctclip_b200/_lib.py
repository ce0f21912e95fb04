"""ctypes binding of the C ABI declared in include/ctclip_b200.h.

PyTorch is used for device memory and streams only; every numeric operation on the product path
is one of the hand-written sm_100a kernels behind these entry points.  There is no CPU path and
no fallback: a missing library or a non-sm_100 device raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("CTCLIP_B200_LIB", _HERE / "libctclip_b200.so"))

P, I, L, F = c_void_p, c_int, c_int64, c_float

# name -> argument ctypes (all functions return int status, except the two noted below)
SIGNATURES = {
    "ctc_device_check": [],
    "ctc_gemm_bf16": [P, L, P, L, P, L, I, I, I, I, P, P, L, P, L, I, P],
    "ctc_gemm_row_perm": [P],
    "ctc_patchify_ln_fwd": [P, L, I, I, I, I, I, I, P, P, F, P, P, F, P, P],
    "ctc_patchify_ln_bwd": [P, L, I, I, I, I, I, I, P, F, P, P, P, I, F, P],
    "ctc_layernorm_fwd": [P, I, I, P, P, F, P, P, P, P],
    "ctc_layernorm_bwd": [P, P, I, I, P, F, P, I, P, P],
    "ctc_layernorm_bwd_bf16": [P, P, I, I, P, F, P, I, P, P],
    "ctc_peg": [P, I, I, I, I, I, P, P, I, I, P, P, P],
    "ctc_peg_frames": [P, P, P, I, I, I, I, P, P, P, P],
    "ctc_frames_gather": [P, P, P, I, L, P, P],
    "ctc_rows_fill": [P, P, I, I, P, P],
    "ctc_attention_fwd": [P, L, P, P, L, I, I, I, I, I, P, P, F, P, I, P, P, P],
    "ctc_attention_fwd_tc": [P, L, P, P, L, I, I, I, I, I, P, P, F, P, F, P, P, P],
    "ctc_attention_score_bound": [P, P, F, P, I, I, I, P, P],
    "ctc_attention_bwd": [P, L, P, P, L, P, P, P, I, I, I, I, I, P, P, F, P, I, P, L, P, P, L, P, P],
    "ctc_attention_probs": [P, L, P, L, P, I, I, I, I, I, P, P, F, P, I, P, P],
    "ctc_attention_fused_probs": [P, L, P, L, P, I, I, I, I, I, P, P, F, P, I, I, P, P, P, P],
    "ctc_geglu_fwd": [P, I, I, P, P],
    "ctc_geglu_bwd": [P, P, I, I, P, P],
    "ctc_cpb_table": [P, P, P, P, P, P, I, I, I, I, P, P],
    "ctc_vq_argmax": [P, P, I, I, P, P, I, P, P, P, P],
    "ctc_vq_gather_pool": [P, P, I, I, I, I, P, P, P, P],
    "ctc_vq_bwd": [P, P, P, I, I, I, I, I, P, P],
    "ctc_latent_proj": [P, P, P, I, L, I, P, I, P, P],
    "ctc_latent_proj_bwd": [P, P, I, L, I, P, P],
    "ctc_text_latent": [P, P, I, I, I, P, P],
    "ctc_latent_sim": [P, P, I, I, I, F, P, P, P, P],
    "ctc_latent_sim_bwd": [P, P, P, I, I, I, F, P, P],
    "ctc_rollout_spatial": [P, I, I, I, P, P],
    "ctc_rollout_temporal": [P, I, I, I, I, P, P],
    "ctc_attn_colmean": [P, I, I, I, P, P],
    "ctc_rollout_fuse": [P, I, I, I, I, I, P, P],
    "ctc_matmul_f32": [P, P, I, P, P],
    "ctc_batch_sum": [P, I, L, F, I, P, P],
    "ctc_colmean": [P, I, I, P, P, P],
    "ctc_gradcam": [P, P, P, I, I, P, P],
    "ctc_upsample_trilinear": [P, I, I, I, P, I, I, I, I, P],
    "ctc_ig_combine": [P, P, L, F, P, P, P],
    "ctc_preprocess_ct": [P, I, I, I, I, L, L, L, F, F, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                          ctypes.c_double, I, I, I, F, P, P, P],
    "ctc_minmax": [P, L, P, P],
    "ctc_patch_is_constant": [P, I, I, I, I, I, F, P, P],
    "ctc_pair_softmax": [P, I, I, P, P],
    "ctc_normalize": [P, I, I, I, P, I, I, P, P],
    "ctc_hist16": [P, L, I, ctypes.c_uint, P, P],
    "ctc_ig_finalize": [P, I, I, I, F, F, F, F, I, P, P],
    "ctc_kth_value": [P, L, L, P, P, P],
    "ctc_quantile_lerp": [P, P, F, P, P],
    "ctc_ig_finalize_dev": [P, I, I, I, P, P, I, P, P],
    "ctc_occlusion_heatmap": [P, P, I, I, I, I, I, I, I, I, I, I, I, I, P, P],
}
OTHER_SYMBOLS = {"ctc_version": (c_int, []), "ctc_last_error": (c_char_p, []),
                 "ctc_launch_count": (ctypes.c_longlong, []), "ctc_vq_num_candidates": (c_int, [c_int]),
                 "ctc_attention_set_tc_bwd": (c_int, [c_int]), "ctc_attention_set_exp2_poly": (c_int, [c_int]), "ctc_colmean_ws_floats": (c_int, [c_int, c_int]),
                 "ctc_kth_value_ws_bytes": (c_int, [])}

EPI_BF16, EPI_F32, EPI_ARGMAX, EPI_GEGLU, EPI_GEGLU_BWD = 0, 1, 2, 3, 4
GEMM_TCGEN05, GEMM_SIMT, GEMM_TCGEN05_1CTA, GEMM_TCGEN05_PAIR = 0, 1, 2, 3
GEMM_BPERM = 0x100   # flag OR-ed into impl: B rows permuted (gemm_row_perm) for the staging-free direct epilogues


def gemm_row_perm():
    """perm[a] = output channel (inside its group of 32) whose weights go to row a of that group of B."""
    arr = (c_int * 32)()
    if load().ctc_gemm_row_perm(arr) != 0:
        raise RuntimeError(load().ctc_last_error().decode())
    return list(arr)
MODE_SPATIAL, MODE_TEMPORAL = 0, 1

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (no GPU needed for this step)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"ctclip_b200: CUDA library {LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (nvcc, sm_100a). There is no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = c_int
        for name, (res, args) in OTHER_SYMBOLS.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = res
        # A/B switches of the measurement tools (defaults are the fastest measured variants)
        if os.environ.get("CTC_ATTN_BWD"):
            lib.ctc_attention_set_tc_bwd(int(os.environ["CTC_ATTN_BWD"]))
        if os.environ.get("CTC_ATTN_EXP2_POLY"):
            lib.ctc_attention_set_exp2_poly(int(os.environ["CTC_ATTN_EXP2_POLY"]))
        _lib = lib
    return _lib


def require_device() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("ctclip_b200: no CUDA device — this package is sm_100a (B200) only, no CPU fallback")
    call("ctc_device_check")


def ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        if not t.is_cuda:
            raise RuntimeError("ctclip_b200: expected a CUDA tensor")
        return t.data_ptr()
    return t


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args) -> None:
    """Invoke one C-ABI entry point; raises RuntimeError with ctc_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*[ptr(a) for a in args])
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.ctc_last_error().decode(errors='replace')}")


def vq_num_candidates(codebook_size: int) -> int:
    """Width of the candidate scratch ctc_vq_argmax expects (include/ctclip_b200.h)."""
    return int(load().ctc_vq_num_candidates(int(codebook_size)))


def launch_count() -> int:
    """Number of kernels the library has launched in this process (bench.py's gpu_launches)."""
    return int(load().ctc_launch_count())
