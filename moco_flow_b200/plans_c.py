"""ctypes view of the C-ABI layer-program builders (mcf_plan_forward / _backward / _gradients, csrc/plan_host.cu).

The Python shim builds its programs with ``plans.py``; these wrappers exist for hosts that want the tables from the
library itself and for ``tests/test_host_cpu.py``, which holds the two builders to each other entry by entry.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from . import _lib as L

MAX_PACK, MAX_CHUNKS, MAX_ROUNDS, MAX_LAYERS, MAX_JOBS, MAX_TENSORS = 192, 128, 24, 17, 32, 32
_i32, _u32 = C.c_int32, C.c_uint32


class PlanSpec(C.Structure):
    _fields_ = [("family", _i32), ("D", _i32), ("W", _i32), ("cx", _i32), ("n_skips", _i32), ("skips", _i32 * 8),
                ("extra_dim", _i32), ("use_quat", _i32), ("sigma_only", _i32), ("training", _i32), ("need_dx", _i32),
                ("nof_kernel", _i32), ("no_pair_merge", _i32)]


class CPlan(C.Structure):
    _fields_ = [("n_pack", _i32), ("n_chunks", _i32), ("n_rounds", _i32), ("n_tensors", _i32),
                ("pack", C.c_uint8 * (L.PACK_DT.itemsize * MAX_PACK)),
                ("chunks", C.c_uint8 * (L.CHUNK_DT.itemsize * MAX_CHUNKS)),
                ("rounds", C.c_uint8 * (L.ROUND_DT.itemsize * MAX_ROUNDS)),
                ("tensor_ids", _i32 * MAX_TENSORS),
                ("wpack_bytes", _u32), ("n_consts", _u32), ("save_tile_bytes", _u32), ("mask_tile_words", _u32),
                ("n_raybias", _i32), ("kind", _i32), ("resident", _i32), ("width", _i32),
                ("save_x0", _u32), ("save_feat", _u32), ("save_he", _u32), ("mask_he", _u32),
                ("save_h", _u32 * MAX_LAYERS), ("mask_h", _u32 * MAX_LAYERS),
                ("save_dhead", _u32), ("save_dye", _u32), ("save_dyf", _u32), ("save_ghead", _u32),
                ("save_dy", _u32 * MAX_LAYERS)]

    def table(self, name: str, dt: np.dtype, n: int) -> np.ndarray:
        return np.frombuffer(bytes(getattr(self, name)), dtype=dt)[:n].copy()


class CGradPlan(C.Structure):
    _fields_ = [("n_jobs", _i32), ("n_unpack", _i32), ("n_params", _i32),
                ("jobs", C.c_uint8 * (L.DWJOB_DT.itemsize * MAX_JOBS)),
                ("job_params", (_i32 * 2) * MAX_JOBS),
                ("unpack", C.c_uint8 * (L.UNPACK_DT.itemsize * L.MAX_UNPACK_PTRS)),
                ("unpack_param", _i32 * L.MAX_UNPACK_PTRS), ("unpack_inner", _u32 * L.MAX_UNPACK_PTRS),
                ("staging_floats", _u32), ("head_ncols", _i32), ("head_stride", _i32), ("head_off", _u32),
                ("param_offset", _u32 * MAX_TENSORS), ("total_floats", _u32)]


def spec(family: int, D: int, W: int, cx: int, skips: Sequence[int], extra_dim: int, use_quat: bool = False,
         sigma_only: bool = False, training: bool = False, need_dx: bool = False, nof_kernel: int = 2,
         pair_merge: bool = True) -> PlanSpec:
    s = PlanSpec()
    s.family, s.D, s.W, s.cx, s.extra_dim = family, D, W, cx, extra_dim
    s.n_skips = len(skips)
    for i, k in enumerate(skips):
        s.skips[i] = k
    s.use_quat, s.sigma_only, s.training, s.need_dx, s.nof_kernel = int(use_quat), int(sigma_only), int(training), \
        int(need_dx), nof_kernel
    s.no_pair_merge = int(not pair_merge)
    return s


def parameter_names(family: int, D: int) -> List[str]:
    """Canonical parameter ids -> reference state_dict names."""
    pre = "xyz_encoding" if family == 0 else "nof_encoding"
    names = []
    for i in range(D):
        names += [f"{pre}_{i+1}.0.weight", f"{pre}_{i+1}.0.bias"]
    names += [f"{pre}_final.weight", f"{pre}_final.bias"]
    if family == 0:
        names += ["extra_encoding.0.weight", "extra_encoding.0.bias", "sigma.weight", "sigma.bias", "rgb.0.weight",
                  "rgb.0.bias"]
    return names


def forward(s: PlanSpec) -> CPlan:
    out = CPlan()
    L.check_rc(L.lib().mcf_plan_forward(C.byref(s), C.byref(out)), "mcf_plan_forward")
    return out


def backward(s: PlanSpec, fwd: CPlan) -> CPlan:
    out = CPlan()
    L.check_rc(L.lib().mcf_plan_backward(C.byref(s), C.byref(fwd), C.byref(out)), "mcf_plan_backward")
    return out


def gradients(s: PlanSpec, fwd: CPlan, bwd: CPlan) -> CGradPlan:
    out = CGradPlan()
    L.check_rc(L.lib().mcf_plan_gradients(C.byref(s), C.byref(fwd), C.byref(bwd), C.byref(out)), "mcf_plan_gradients")
    return out
