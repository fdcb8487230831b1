"""ctypes binding of the mst C-ABI (include/mst.h) and the in-tree build recipe.

The shared library is the product: every device operation of the hot path
goes through it.  There is no fallback - if the library is missing or a call
fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG_DIR, "csrc")
LIB_PATH = os.path.join(_PKG_DIR, "libmst_b200.so")
SOURCES = ["api.cu", "update.cu", "simt.cu", "tc_gemm.cu", "tc_attn.cu", "train.cu", "text.cu"]
HEADERS = ["common.cuh", "simt.cuh", "tc.cuh", "tc_ptx.cuh", "smem_gemm.cuh", "dropout.cuh", os.path.join("..", "..", "include", "mst.h")]

MAX_LAYERS = 32
PREC_FP32, PREC_BF16 = 0, 1
SAMPLER_DDPM, SAMPLER_DDIM = 0, 1
NOISE_NONE, NOISE_TENSOR, NOISE_PHILOX = 0, 1, 2
MASK_NONE, MASK_FULL, MASK_FT, MASK_F = 0, 1, 2, 3

# every symbol include/mst.h declares
EXPORTED = [
    "mst_version", "mst_last_error", "mst_launch_count", "mst_profile_begin", "mst_profile_end", "mst_profile_collect", "mst_device_info", "mst_abi_sizes", "mst_engine_create", "mst_engine_destroy",
    "mst_engine_packed_weight_bytes", "mst_engine_load_weights", "mst_engine_workspace_bytes", "mst_time_embed",
    "mst_text_embed", "mst_denoiser_forward", "mst_update_step", "mst_q_sample", "mst_cfg_combine",
    "mst_philox_normal", "mst_train_sizes", "mst_denoiser_forward_train", "mst_denoiser_backward", "mst_abi_sizes_train",
    "mst_motion_encoder_forward", "mst_motion_encoder_backward", "mst_masked_l2", "mst_update_step_backward",
    "mst_adamw_step", "mst_sumsq2", "mst_recover_from_ric",
    "mst_clip_text_create", "mst_clip_text_destroy", "mst_abi_sizes_clip_text", "mst_clip_text_packed_weight_bytes",
    "mst_clip_text_workspace_bytes", "mst_clip_text_load_weights", "mst_clip_text_encode", "mst_test_dropout_scale", "mst_test_gemm_bf16", "mst_test_gemm_epi_bf16", "mst_test_set_gemm_debug", "mst_test_attention_bf16",
]


class ModelDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("n_feats", "d_model", "n_heads", "d_ff", "n_layers", "clip_dim", "pe_len", "precision")]


_LAYER_FIELDS = ("qkv_w", "qkv_b", "o_w", "o_b", "w1", "b1", "w2", "b2", "ln1_g", "ln1_b", "ln2_g", "ln2_b")


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _LAYER_FIELDS]


class Weights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("in_w", "in_b", "pe", "t_w1", "t_b1", "t_w2", "t_b2", "txt_w", "txt_b", "out_w", "out_b")] + \
               [("layers", LayerWeights * MAX_LAYERS)]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_frames", C.c_int32), ("cfg", C.c_int32), ("uncond", C.c_int32),
        ("x", C.c_void_p), ("temb", C.c_void_p), ("temb_row_dev", C.c_void_p), ("temb_row_offset", C.c_int32),
        ("text_emb", C.c_void_p), ("out_cond", C.c_void_p), ("out_uncond", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("use_graph", C.c_int32),
        ("dropout_p", C.c_float), ("dropout_seed", C.c_void_p), ("tape_seqs", C.c_int32), ("tape_seq_offset", C.c_int32),
    ]


class UpdateArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_feats", C.c_int32), ("n_frames", C.c_int32),
        ("sampler", C.c_int32), ("clip_denoised", C.c_int32),
        ("out_cond", C.c_void_p), ("out_uncond", C.c_void_p), ("cfg_scale", C.c_void_p),
        ("x_t", C.c_void_p), ("x_prev", C.c_void_p), ("pred_xstart", C.c_void_p),
        ("mask_kind", C.c_int32), ("mask", C.c_void_p), ("x_inpaint", C.c_void_p), ("mask_noise", C.c_int32),
        ("t_vec", C.c_void_p), ("t_scalar_dev", C.c_void_p), ("t_imm", C.c_int32), ("advance_t", C.c_int32),
        ("block_counter", C.c_void_p),
        ("coef1", C.c_void_p), ("coef2", C.c_void_p), ("sigma", C.c_void_p), ("recip", C.c_void_p),
        ("recipm1", C.c_void_p),
        ("noise_kind", C.c_int32), ("noise", C.c_void_p), ("const_noise", C.c_int32),
        ("philox_seed", C.c_uint64), ("philox_sample_offset", C.c_uint64),
    ]


class LayerGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _LAYER_FIELDS]


class BackwardArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_frames", C.c_int32), ("d_out", C.c_void_p), ("d_x", C.c_void_p),
        ("layer_grads", C.POINTER(LayerGrads)), ("tape", C.c_void_p), ("tape_bytes", C.c_size_t),
        ("scratch", C.c_void_p), ("scratch_bytes", C.c_size_t), ("use_graph", C.c_int32),
        ("dropout_p", C.c_float), ("dropout_seed", C.c_void_p), ("tape_seqs", C.c_int32), ("tape_seq_offset", C.c_int32),
    ]


class ClipTextDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("vocab", "ctx", "width", "n_heads", "n_layers", "d_ff", "d_out", "precision")]


_CLIP_LAYER_FIELDS = ("ln1_g", "ln1_b", "qkv_w", "qkv_b", "o_w", "o_b", "ln2_g", "ln2_b", "fc_w", "fc_b", "proj_w", "proj_b")


class ClipTextLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _CLIP_LAYER_FIELDS]


class ClipTextWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("token_embedding", "positional_embedding", "lnf_g", "lnf_b", "text_projection")] + \
               [("layers", ClipTextLayer * MAX_LAYERS)]


def nvcc_command(out_path: str = LIB_PATH) -> list[str]:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("MST_NVCC_FLAGS", "").split()
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"] + extra + [
        "-Xcompiler", "-fPIC", "-shared", "-o", out_path] + [os.path.join(_CSRC, s) for s in SOURCES]


HASH_PATH = LIB_PATH + ".srchash"


def source_hash() -> str:
    """sha256 over the CUDA sources and headers the library is built from (content, not mtimes: the snapshot that
    carries the built .so to a GPU box does not preserve timestamps)."""
    import hashlib
    h = hashlib.sha256()
    for d in sorted([os.path.join(_CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(_CSRC, x)) for x in HEADERS]):
        if os.path.exists(d):
            h.update(os.path.basename(d).encode())
            with open(d, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    """True when the library is missing or was built from other sources than the ones on disk."""
    if not os.path.exists(LIB_PATH):
        return True
    try:
        with open(HASH_PATH) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libmst_b200.so next to this file."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    with open(HASH_PATH, "w") as f:
        f.write(source_hash() + "\n")
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def _declare(lib):
    vp, i32, i64, u64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_size_t
    lib.mst_version.restype = C.c_char_p
    lib.mst_last_error.restype = C.c_char_p
    lib.mst_launch_count.restype = C.c_uint64
    lib.mst_launch_count.argtypes = []
    sigs = {
        "mst_device_info": [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "mst_profile_begin": [vp],
        "mst_profile_end": [C.POINTER(C.c_float), C.c_char_p, i32, sz, C.POINTER(i32)],
        "mst_profile_collect": [C.POINTER(C.c_float), C.c_char_p, i32, sz, C.POINTER(i32)],
        "mst_abi_sizes": [C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)],
        "mst_engine_create": [C.POINTER(ModelDesc), C.POINTER(vp)],
        "mst_engine_destroy": [vp],
        "mst_engine_packed_weight_bytes": [vp, C.POINTER(sz)],
        "mst_engine_load_weights": [vp, C.POINTER(Weights), vp, sz, vp],
        "mst_engine_workspace_bytes": [vp, C.c_int, C.c_int, C.POINTER(sz)],
        "mst_time_embed": [vp, vp, C.c_int, vp, vp, sz, vp],
        "mst_text_embed": [vp, vp, C.c_int, vp, vp],
        "mst_denoiser_forward": [vp, C.POINTER(ForwardArgs), vp],
        "mst_update_step": [C.POINTER(UpdateArgs), vp],
        "mst_q_sample": [vp, vp, i32, vp, vp, i32, vp, vp, vp, i32, i32, i32, vp],
        "mst_cfg_combine": [vp, vp, vp, vp, i32, i64, vp],
        "mst_philox_normal": [vp, i32, i64, u64, u64, i32, vp],
        "mst_train_sizes": [vp, i32, i32, C.POINTER(sz), C.POINTER(sz)],
        "mst_denoiser_forward_train": [vp, C.POINTER(ForwardArgs), vp, sz, vp],
        "mst_denoiser_backward": [vp, C.POINTER(BackwardArgs), vp],
        "mst_abi_sizes_train": [C.POINTER(sz), C.POINTER(sz)],
        "mst_motion_encoder_forward": [vp, vp, vp, vp, vp, i32, i32, vp, vp, sz, C.c_float, vp, i32, vp],
        "mst_motion_encoder_backward": [vp, vp, i32, i32, vp, vp, sz, vp, sz, C.c_float, vp, i32, vp],
        "mst_test_dropout_scale": [vp, i64, C.c_float, vp, i32, vp],
        "mst_masked_l2": [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "mst_update_step_backward": [vp, vp, vp, vp, i32, vp, vp, i32, vp, i32, i32, i32, vp],
        "mst_adamw_step": [vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i32, C.c_float, vp],
        "mst_sumsq2": [vp, vp, i64, vp, vp],
        "mst_recover_from_ric": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "mst_clip_text_create": [vp, vp],
        "mst_clip_text_destroy": [vp],
        "mst_abi_sizes_clip_text": [vp, vp],
        "mst_clip_text_packed_weight_bytes": [vp, vp],
        "mst_clip_text_workspace_bytes": [vp, i32, vp],
        "mst_clip_text_load_weights": [vp, vp, vp, sz, vp],
        "mst_clip_text_encode": [vp, vp, i32, vp, vp, sz, vp],
        "mst_test_gemm_bf16": [vp, vp, vp, vp, i32, i32, i32, vp],
        "mst_test_gemm_epi_bf16": [i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "mst_test_set_gemm_debug": [vp],
        "mst_test_attention_bf16": [vp, vp, vp, i32, i32, vp, sz, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int


def _check_abi(lib):
    """ctypes mirrors vs the structs the library was compiled with: a stale binary or a header edit without a matching
    edit here would otherwise corrupt arguments silently."""
    sz = [C.c_size_t() for _ in range(4)]
    lib.mst_abi_sizes(*[C.byref(v) for v in sz])
    got = [v.value for v in sz]
    want = [C.sizeof(ModelDesc), C.sizeof(Weights), C.sizeof(ForwardArgs), C.sizeof(UpdateArgs)]
    names = ["mst_model_desc", "mst_weights", "mst_forward_args", "mst_update_args"]
    a, b = C.c_size_t(), C.c_size_t()
    lib.mst_abi_sizes_train(C.byref(a), C.byref(b))
    got += [a.value, b.value]
    want += [C.sizeof(LayerGrads), C.sizeof(BackwardArgs)]
    names += ["mst_layer_grads", "mst_backward_args"]
    bad = [f"{n}: library {g} bytes, ctypes {w}" for n, g, w in zip(names, got, want) if g != w]
    if bad:
        raise RuntimeError(f"{LIB_PATH} does not match this Python package (ABI struct sizes differ: " + "; ".join(bad) +
                           "); rebuild with `python __graft_entry__.py`")


def load():
    """dlopen libmst_b200.so.  The library is (re)built first when it is missing or was built from other sources than
    the ones on disk (content hash, see source_hash) and a compiler is available; a stale library that cannot be
    rebuilt raises instead of loading silently.  MST_NO_BUILD=1 never compiles."""
    global _lib
    with _lock:
        if _lib is None:
            if needs_build():
                can_build = os.environ.get("MST_NO_BUILD") != "1" and os.path.exists(nvcc_command()[0])
                if not can_build:
                    what = "is missing" if not os.path.exists(LIB_PATH) else "was built from different sources"
                    raise RuntimeError(f"{LIB_PATH} {what} and cannot be built here; run `python __graft_entry__.py` "
                                       "(there is no CPU fallback)")
                build(force=True)
            lib = C.CDLL(LIB_PATH)
            _declare(lib)
            _check_abi(lib)
            _lib = lib
    return _lib


def check(rc: int, what: str = "mst"):
    if rc != 0:
        msg = load().mst_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
