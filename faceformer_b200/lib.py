"""ctypes binding of libffb200.so (the C ABI declared in include/ffb200.h).

There is NO fallback: if the shared library is missing, or no sm_100 GPU is
visible, every compute entry point raises.  ``build()`` compiles the library
in-tree with nvcc for sm_100a (cross-compiles on a CPU-only box).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libffb200.so")
SOURCES = [os.path.join(HERE, "csrc", "ffb200.cu")]
HEADERS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + [os.path.join(ROOT, "include", "ffb200.h")]

FFB_ABI_VERSION = 1
FFB_HOST, FFB_DEVICE = 0, 1
FFB_OPT_DEDUP_PAD, FFB_OPT_PRUNE_LAST, FFB_OPT_TIMING, FFB_OPT_PROFILE, FFB_OPT_TENSOR_CORE, FFB_OPT_ATTN_MMA, FFB_OPT_TC_FORMAT, FFB_OPT_STAGGER, FFB_OPT_TMA_EPILOGUE, FFB_OPT_ATTN_X, FFB_OPT_GEMM_VARIANT, FFB_OPT_ENCODER_TC, FFB_OPT_PDL, FFB_OPT_POINTER_BATCHED, FFB_OPT_BEAM, FFB_OPT_ENCODER_PRECISION, FFB_OPT_HEAD_FP64, FFB_OPT_FORCE_F, FFB_OPT_ENCODE_ONLY, FFB_OPT_ATTN_LONG, FFB_OPT_SKINNY_GEMM, FFB_OPT_L0_CACHE, FFB_OPT_PERSISTENT = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23
PROFILE_CLASSES = ("linear", "layernorm", "attn_rows", "attn_tiled", "pointer", "other", "linear_tc")


class FFBError(RuntimeError):
    """Raised for any non-zero status of the C ABI (message from ffb_last_error)."""


class ffb_config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "mode", "num_model", "num_head", "num_feedforward", "num_encoder_layers",
        "num_decoder_layers", "in_dim", "num_lines", "num_token", "seq_len", "device")]


# symbol -> (restype, argtypes); every function include/ffb200.h declares
_P = C.c_void_p
SIGNATURES = {
    "ffb_weight_count": (C.c_size_t, [C.POINTER(ffb_config)]),
    "ffb_create": (C.c_int, [C.POINTER(ffb_config), C.POINTER(_P)]),
    "ffb_destroy": (C.c_int, [_P]),
    "ffb_last_error": (C.c_char_p, [_P]),
    "ffb_set_option": (C.c_int, [_P, C.c_int, C.c_int]),
    "ffb_load_weights": (C.c_int, [_P, _P, C.c_size_t, C.c_int, _P]),
    "ffb_encode": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int, _P]),
    "ffb_batch_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ffb_decode_greedy": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_int32), _P]),
    "ffb_forward_eval": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, C.c_int, C.POINTER(C.c_int32), _P]),
    "ffb_featurize": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P, _P, C.c_int, _P]),
    "ffb_parse_faces": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, C.c_double, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "ffb_get_memory": (C.c_int, [_P, _P, C.c_int, _P]),
    "ffb_set_memory": (C.c_int, [_P, _P, C.c_int, _P]),
    "ffb_get_last_logits": (C.c_int, [_P, _P, C.c_int, _P]),
    "ffb_get_last_pointer": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.c_int, _P]),
    "ffb_forced_prefix_logits": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int, _P]),
    "ffb_kernel_launches": (C.c_int64, [_P]),
    "ffb_fp16_fallbacks": (C.c_int, [_P]),
    "ffb_stop_exchange_export": (C.c_int, [_P, _P]),
    "ffb_stop_exchange_connect": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "ffb_stop_exchange_disconnect": (C.c_int, [_P]),
    "ffb_get_beams": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "ffb_overflowed": (C.c_int, [_P, C.POINTER(C.c_int32), _P]),
    "ffb_steps_launched": (C.c_int, [_P]),
    "ffb_used_persistent": (C.c_int, [_P]),
    "ffb_forward_train": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P, C.c_int32, _P, _P, C.c_int32, _P]),
    "ffb_phase_times": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int32]),
    "ffb_profile_read": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "ffb_op_linear": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "ffb_op_linear_tc": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "ffb_bench_linear_tc": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), _P]),
    "ffb_op_layernorm": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "ffb_op_attention": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, _P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
}

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-shared", "-Xcompiler", "-fPIC"]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libffb200.so in-tree for sm_100a.  Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise FFBError("nvcc not found: cannot build libffb200.so (there is no CPU fallback)")
    extra = os.environ.get("FFB_NVCC_EXTRA", "").split()          # tuning experiments only (e.g. -DFFB_DRAIN_KB=8)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise FFBError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """dlopen libffb200.so and bind every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FFBError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(faceformer_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error(handle) -> str:
    msg = load().ffb_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, handle=None) -> None:
    if status != 0:
        raise FFBError(f"libffb200 status {status}: {last_error(handle)}")
