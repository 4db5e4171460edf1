"""Build and load ``libmms2ut_b200.so`` (the C-ABI kernel library) with ctypes.

The library is compiled IN-TREE by nvcc for sm_100a only (``build()``); there is no JIT cache, no
torch C++ extension and no CPU fallback: if the shared object is missing or fails to load, every op
raises.  Signatures mirror ``include/mms2ut_b200.h``.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path
from typing import List, Optional

_PKG = Path(__file__).resolve().parent
CSRC = _PKG / "csrc"
INCLUDE = _PKG.parent / "include"
LIB_PATH = _PKG / "libmms2ut_b200.so"
SOURCES = ["abi.cu", "gemm.cu", "gemm_ln.cu", "rowwise.cu", "fbank.cu", "attention.cu", "cross_attention.cu", "attention_bwd.cu", "backward.cu", "wgrad.cu", "heads_gemm.cu", "attention_bwd_fused.cu", "p2p.cu", "gemm_ln_bwd.cu"]
ABI_VERSION = 5

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libmms2ut_b200.so")


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + list(CSRC.glob("*.cuh")) + [INCLUDE / "mms2ut_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu for sm_100a and link the shared library (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + os.environ.get("MM_NVCC_EXTRA", "").split()
    objdir = _PKG / "build"
    objdir.mkdir(exist_ok=True)
    procs = []
    for s in SOURCES:
        cmd = [nvcc, *flags, "-I", str(INCLUDE), "-c", str(CSRC / s), "-o", str(objdir / (s[:-3] + ".o"))]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
    objs = [str(objdir / (s[:-3] + ".o")) for s in SOURCES]
    cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


class GemmArgs(C.Structure):
    """``mm_gemm_args`` (include/mms2ut_b200.h)."""

    _fields_ = [
        ("a0", C.c_void_p), ("a1", C.c_void_p), ("w", C.c_void_p),
        ("a0_ld", C.c_int64), ("a0_bs", C.c_int64), ("a1_ld", C.c_int64), ("a1_bs", C.c_int64),
        ("w_ld", C.c_int64), ("w_bs", C.c_int64),
        ("rows", C.c_int32), ("batches", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("k_split", C.c_int32), ("w_batched", C.c_int32),
        ("dtype", C.c_int32), ("mode", C.c_int32), ("block_n", C.c_int32),
        ("bias", C.c_void_p), ("scale", C.c_float), ("scale_cols", C.c_int32),
        ("out0", C.c_void_p), ("out0_ld", C.c_int64), ("out0_bs", C.c_int64),
        ("out1", C.c_void_p), ("out1_ld", C.c_int64), ("out1_bs", C.c_int64),
        ("aux0", C.c_void_p), ("aux1", C.c_void_p), ("aux_ld", C.c_int64),
        ("rows_per_seq", C.c_int32), ("out_tbc", C.c_int32), ("n_seqs", C.c_int32), ("out_row_offset", C.c_int32),
        ("vt", C.c_void_p), ("vt_col0", C.c_int32), ("vt_rows", C.c_int32), ("vt_ld", C.c_int64),
        ("pos", C.c_void_p), ("seq_lens", C.c_void_p),
        ("a_mn", C.c_int32), ("w_mn", C.c_int32), ("a_kbatch", C.c_int32), ("w_kbatch", C.c_int32),
        ("a_hm", C.c_int32), ("w_hm", C.c_int32), ("out_hm", C.c_int32), ("heads", C.c_int32),
        ("head_stride", C.c_int32), ("a_k_total", C.c_int64), ("w_k_total", C.c_int64),
        ("drop_p", C.c_float), ("drop_site", C.c_uint32), ("drop_seed", C.c_uint64), ("drop_seed_dev", C.c_void_p),
    ]


class ReduceJob(C.Structure):
    """``mm_reduce_job`` (include/mms2ut_b200.h)."""

    _fields_ = [("part", C.c_void_p), ("out", C.c_void_p), ("stride", C.c_int64), ("n", C.c_int64),
                ("n_partials", C.c_int32), ("accumulate", C.c_int32)]


class WgradGroup(C.Structure):
    """``mm_wgrad_group`` (include/mms2ut_b200.h)."""

    _fields_ = [("dy", C.c_void_p), ("x", C.c_void_p), ("out", C.c_void_p), ("dy_ld", C.c_int64), ("x_ld", C.c_int64),
                ("out_ld", C.c_int64), ("n_out", C.c_int32), ("k_in", C.c_int32), ("bias", C.c_void_p),
                ("tokens", C.c_int64)]


WGRAD_MAX_GROUPS = 64

EXPORTS = {
    # name: (restype, argtypes)
    "mm_abi_version": (C.c_int, []),
    "mm_last_error": (C.c_char_p, []),
    "mm_fbank_table_floats": (C.c_int, []),
    "mm_fbank_build_tables": (C.c_int, [C.c_void_p]),
    "mm_fbank_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                               C.c_void_p]),
    "mm_fbank_i16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                               C.c_void_p]),
    "mm_cmvn_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mm_cmvn_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "mm_cmvn_apply_specaug": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                        C.c_float, C.c_void_p]),
    "mm_seq_lens": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mm_seq_lens_mask": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mm_gemm": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "mm_padding_mask": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mm_label_smoothed_nll": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "mm_embed_tokens": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mm_attention": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                               C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                               C.c_int64, C.c_int32, C.c_void_p]),
    "mm_self_attention_lse": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "mm_attention_lse": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                   C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                   C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "mm_attention_bwd_scores": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                          C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "mm_cross_attention": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                     C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                     C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "mm_layernorm_gather": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.c_void_p]),
    "mm_gemm_resid_ln": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                   C.c_int32, C.c_void_p]),
    "mm_gemm_resid_ln_drop": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_float, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p]),
    "mm_gemm_resid_ln_out": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_void_p]),
    "mm_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                               C.c_int32, C.c_float, C.c_void_p]),
    "mm_self_attention": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_int64, C.c_int32, C.c_void_p]),
    "mm_softmax_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                  C.c_int64, C.c_int32, C.c_void_p]),
    "mm_mask_scores": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "mm_convert_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    # training-step variant (csrc/backward.cu)
    "mm_pack_t": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                            C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                            C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "mm_rowsum": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "mm_reduce_partials": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "mm_reduce_partials_many": (C.c_int, [C.POINTER(ReduceJob), C.c_int32, C.c_void_p]),
    "mm_attention_bwd_fused": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "mm_ipc_get_handle": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mm_ipc_open_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mm_ipc_close_handle": (C.c_int, [C.c_void_p]),
    "mm_p2p_barrier": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "mm_p2p_allreduce_f32": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "mm_gemm_ln_bwd_partial_rows": (C.c_int, [C.c_int64]),
    "mm_gemm_ln_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_uint64, C.c_void_p,
                                 C.c_uint32, C.c_int32, C.c_void_p]),
    "mm_p2p_pack_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "mm_p2p_allreduce_bf16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "mm_p2p_unpack_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "mm_attention_bwd_general_scratch_floats": (C.c_int64, [C.c_int32]),
    "mm_attention_bwd_general": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                           C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                           C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "mm_attention_drop": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                    C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_int64, C.c_void_p, C.c_float, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int32,
                                    C.c_void_p]),
    "mm_attention_bwd_general_drop": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                                C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                                C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                                C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                                C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_float, C.c_uint64,
                                                C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p]),
    "mm_self_attention_drop": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_float, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int32,
                                         C.c_void_p]),
    "mm_attention_bwd_fused_drop": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                              C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_int64, C.c_float, C.c_uint64, C.c_void_p, C.c_uint32,
                                              C.c_int32, C.c_void_p]),
    "mm_heads_gemm": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                C.c_float, C.c_int32, C.c_void_p]),
    "mm_wgrad_grouped": (C.c_int, [C.POINTER(WgradGroup), C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "mm_layernorm_bwd_blocks": (C.c_int, []),
    "mm_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mm_layernorm_bwd_drop": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_uint64, C.c_void_p, C.c_uint32,
                                        C.c_int32, C.c_void_p]),
    "mm_softmax_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "mm_softmax_dropout_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                         C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                         C.c_int32, C.c_float, C.c_uint64, C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p]),
    "mm_label_smoothed_nll_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_float,
                                            C.c_float, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "mm_embed_tokens_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p,
                                      C.c_int32, C.c_void_p]),
    "mm_colsum_blocks": (C.c_int, [C.c_int32]),
    "mm_colsum": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                            C.c_void_p]),
    "mm_glu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_int32, C.c_void_p]),
    "mm_gate_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                              C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mm_tbc_to_btc": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mm_col2im_k5s2": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "mm_dropout": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_uint64, C.c_void_p,
                             C.c_uint32, C.c_int32, C.c_void_p]),
    "mm_sumsq_blocks": (C.c_int, []),
    "mm_grad_clip_coef": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_void_p]),
    "mm_adam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float,
                          C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the in-tree library and bind every symbol the header declares.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    v = lib.mm_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"libmms2ut_b200.so ABI {v} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def header_symbols() -> List[str]:
    """Function names declared in include/mms2ut_b200.h (used by the CPU-side export test)."""
    import re

    txt = (INCLUDE / "mms2ut_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mm_[a-z0-9_]+)\s*\(", txt)))


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mm_last_error().decode(errors="replace")
        raise RuntimeError(f"libmms2ut_b200 {what} failed (code {rc}): {msg}")
