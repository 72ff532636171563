"""ctypes binding of csrc/libmoco_flow_b200.so (the C ABI declared in include/moco_flow_b200.h).

There is no CPU fallback anywhere in this package: if the library is missing or a call fails the
caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCF_LIB_PATH") or os.path.join(_HERE, "csrc", "libmoco_flow_b200.so")  # override: kernel experiments only

MAX_FREQS = 24
TILE_ROWS = 128
BLOCK_BYTES = 16384
NONE = 0xFFFFFFFF

ACT = {"relu": 0, "softplus": 1}

EPI_RELU, EPI_RELU_SIGMA, EPI_LINEAR, EPI_NERF_RGB, EPI_NOF_HEAD = 0, 1, 2, 3, 4
EPI_B_MASK, EPI_B_MASK_SIGMA, EPI_B_LINEAR, EPI_B_DPE = 16, 17, 18, 19
PRO_PE_XYZ, PRO_DENSE, PRO_B_NERF, PRO_B_NOF = 0, 1, 2, 3

CHUNK_DT = np.dtype([("src_off", "<u4"), ("bytes", "<u4"), ("a_buf", "u1"), ("a_kblock", "u1"), ("ksteps", "u1"),
                     ("flags", "u1"), ("n", "<u2"), ("acc_col", "<u2")])
ROUND_DT = np.dtype([("epi", "<u2"), ("n_out", "<u2"), ("acc_col", "<u2"), ("chunk_begin", "<u2"),
                     ("chunk_end", "<u2"), ("raybias", "<i2"), ("const_off", "<u4"), ("aux_off", "<u4"),
                     ("save_off", "<u4"), ("mask_off", "<u4"), ("reserved", "<u4")])
PACK_DT = np.dtype([("dst_off", "<u4"), ("bytes", "<u4"), ("kind", "<i4"), ("tensor", "<i4"), ("row0", "<i4"),
                    ("nrows", "<i4"), ("col0", "<i4"), ("ncols", "<i4"), ("ld", "<i4"), ("transposed", "<i4")])
UNPACK_DT = np.dtype([("src_off", "<u4"), ("dst_off", "<u4"), ("src_ld", "<i4"), ("dst_ld", "<i4"), ("nrows", "<i4"),
                      ("ncols", "<i4"), ("transposed", "<i4"), ("reserved", "<i4")])
DWJOB_DT = np.dtype([("p_off", "<u4"), ("q_off", "<u4"), ("p_src", "<i4"), ("q_src", "<i4"), ("p_cols", "<i4"),
                     ("q_cols", "<i4"), ("st_off", "<u4"), ("ld", "<i4"), ("n_i", "<i4"), ("n_j", "<i4"),
                     ("colsum_off", "<i4"), ("enabled", "<i4"), ("q_split", "<i4"), ("q2_off", "<u4")])
MAX_UNPACK_PTRS = 40
assert DWJOB_DT.itemsize == 56
assert CHUNK_DT.itemsize == 16 and ROUND_DT.itemsize == 32 and PACK_DT.itemsize == 40 and UNPACK_DT.itemsize == 32

_vp, _i32, _i64, _u32, _f32 = C.c_void_p, C.c_int32, C.c_longlong, C.c_uint32, C.c_float


class ChainParams(C.Structure):
    _fields_ = [
        ("chunks", _vp), ("rounds", _vp), ("n_chunks", _i32), ("n_rounds", _i32), ("width", _i32),
        ("prologue", _i32), ("wpack", _vp), ("consts", _vp), ("raybias", _vp * 4),
        ("n_rows", _i64), ("rows_per_ray", _i32), ("n_rays", _i32),
        ("xyz", _vp), ("dense", _vp), ("dense_stride", _i32), ("dense_cols", _i32),
        ("pe_n_freqs", _i32), ("pe_pad_to", _i32), ("pe_freq", _f32 * MAX_FREQS), ("pe_weight", _f32 * MAX_FREQS),
        ("out", _vp), ("out_stride", _i32), ("sigma_col", _i32), ("use_quat", _i32), ("head_save", _vp),
        ("save", _vp), ("save_tile_bytes", _i64), ("masks", _vp), ("mask_tile_words", _i64), ("x0_save_off", _u32),
        ("g_out", _vp), ("fwd_out", _vp), ("fwd_save", _vp), ("fwd_save_tile_bytes", _i64), ("fwd_masks", _vp),
        ("fwd_mask_tile_words", _i64), ("fwd_x0_off", _u32), ("fwd_he_off", _u32),
        ("d_xyz", _vp), ("d_head", _vp), ("rayfeat", _vp), ("rayfeat_stride", _i32), ("rayfeat_dim", _i32),
        ("extra_save_off", _u32), ("dhead_save_off", _u32), ("max_ctas", _i32),
        ("timing", _vp), ("cta_pair", _i32), ("program_kind", _i32), ("pe_table", _vp),
        ("wpack_bytes", _u32), ("resident", _i32), ("d_dense", _vp), ("d_dense_stride", _i32), ("reserved0", _i32),
    ]


class DwParams(C.Structure):
    _fields_ = [
        ("p_base", _vp), ("p_tile_bytes", _i64), ("p_off", _u32), ("p_cols", _i32),
        ("q_base", _vp), ("q_tile_bytes", _i64), ("q_off", _u32), ("q_cols", _i32),
        ("out", _vp), ("ld_out", _i32), ("n_i", _i32), ("n_j", _i32),
        ("colsum_p", _vp), ("n_tiles", _i64), ("max_ctas", _i32),
    ]


EXPORTS = [
    "mcf_abi_version", "mcf_device_error_flag", "mcf_coarse_samples", "mcf_ray_points", "mcf_pe_fwd", "mcf_pe_bwd",
    "mcf_ray_bias", "mcf_composite_fwd", "mcf_composite_bwd", "mcf_sample_pdf", "mcf_masked_l1_fwd",
    "mcf_masked_l1_finalize", "mcf_masked_l1_bwd", "mcf_pack", "mcf_chain_launch", "mcf_dw_gemm", "mcf_dw_gemm_batch", "mcf_rayfeat_image", "mcf_unpack",
    "mcf_unpack_accumulate", "mcf_colsum", "mcf_adam_step", "mcf_make_rays", "mcf_canvas_scatter", "mcf_nearest_vertex",
    "mcf_plan_forward", "mcf_plan_backward", "mcf_plan_gradients",
]

_lib = None


class MocoFlowLibraryError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built (python -m moco_flow_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MocoFlowLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m moco_flow_b200.build` (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        for name in EXPORTS:
            getattr(_lib, name).restype = C.c_int
        if _lib.mcf_abi_version() != 1:
            raise MocoFlowLibraryError("ABI version mismatch; rebuild the library")
    return _lib


# ---- instrumentation used by bench.py ----------------------------------------------------------------
LAUNCHES = 0      # C-ABI kernel-launching calls issued so far (every entry except the two query calls)
PROFILE = None    # when a list: (tag, work, unit, start_event, end_event) per timed launch


class timed:
    """Brackets one launch with CUDA events on the current stream when profiling is switched on."""

    def __init__(self, tag: str, work: float, unit: str):
        self.tag, self.work, self.unit = tag, work, unit

    def __enter__(self):
        if PROFILE is not None:
            import torch
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.ev[1].record()
            PROFILE.append((self.tag, self.work, self.unit, self.ev[0], self.ev[1]))
        return False


def check(code: int, what: str) -> None:
    global LAUNCHES
    LAUNCHES += 1
    if code != 0:
        kind = "cudaError" if code > 0 else "MCF_ERR"
        raise MocoFlowLibraryError(f"{what} failed: {kind} {code}")


def check_rc(code: int, what: str) -> None:
    """Return-code check of an entry that launches nothing (not counted as a GPU launch)."""
    if code != 0:
        raise MocoFlowLibraryError(f"{what} failed: {'cudaError' if code > 0 else 'MCF_ERR'} {code}")


def ptr(t) -> C.c_void_p:
    """Device pointer of a torch tensor (None -> NULL)."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def f32_array(vals, n=None):
    vals = [float(v) for v in vals]
    n = len(vals) if n is None else n
    arr = (C.c_float * max(n, 1))()
    for i, v in enumerate(vals):
        arr[i] = v
    return arr


def device_error_flag() -> int:
    flag = C.c_uint(0)
    lib().mcf_device_error_flag(C.byref(flag))
    return int(flag.value)
