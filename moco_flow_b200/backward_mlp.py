"""Autograd wiring of the fused chains: training forward (saves bf16 operand images + ReLU bit masks),
backward dX chain, weight-gradient GEMMs and the scatter into per-parameter gradients.

Gradient flow follows the reference's autograd graph: parameters of the module and
the input points ``xyz`` (when they come out of a NoF); never the per-ray features or the rays.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from . import ops
from . import plans as P
from .mlp import fold_bias, setup_input


def _n_tiles(M: int) -> int:
    return (M + L.TILE_ROWS - 1) // L.TILE_ROWS


def _dev_table(arr: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(device)


class _TrainState:
    """Cached per-module backward programs (dX plan with/without input gradient, grad job tables)."""

    def __init__(self):
        self.bwd: Dict[bool, ops.PackedPlan] = {}
        self.grad: Dict[bool, P.GradPlan] = {}
        self.unpack_dev: Dict[bool, torch.Tensor] = {}
        self.jobs_dev: Dict[tuple, torch.Tensor] = {}


def _train_state(model) -> _TrainState:
    st = model.__dict__.get("_mcf_train")
    if st is None or model.__dict__.get("_mcf_plans") is None:
        st = _TrainState()
        model.__dict__["_mcf_train"] = st
    return st


# When True, parameter gradients are ADDED into existing ``param.grad`` tensors by the scatter kernel (one launch
# per MLP evaluation) and autograd receives None for them, instead of one autograd accumulation kernel per
# parameter per evaluation (~144 per training step).  Equivalent for ``loss.backward()``; not meaningful for
# ``torch.autograd.grad``.  Switched on by ``dp.FlatGradients(..., fused_accumulate=True)``.
ACCUMULATE_INTO_GRAD = False
# CTAs each weight-gradient job is split over (split-K).  Half the SMs: two jobs of a launch are resident at a time,
# so one job's TMEM drain + fp32 atomics overlap the other's HBM streaming, and every CTA amortises its set-up over
# twice the tiles.  Measured on B200 (4096-ray step): 148 -> 4.27 ms, 74 -> 3.63 ms (0.87 of the HBM copy peak).
_DW_CTAS_ENV = int(__import__('os').environ.get('MCF_DW_CTAS', '0'))


def dw_ctas_per_job(device) -> int:
    if _DW_CTAS_ENV > 0:
        return _DW_CTAS_ENV
    if DW_OVERLAP_SMS > 0:
        return max(2, DW_OVERLAP_SMS // 2)
    return max(2, torch.cuda.get_device_properties(device).multi_processor_count // 2)


# Overlap of the weight-gradient GEMMs with the backward dX chains.  The dX chain of the next MLP evaluation does not
# depend on the weight gradients of the previous one; the GEMMs are HBM-bound (they do not need every SM to saturate
# the memory system) and the chains are SM-bound (they leave most of the HBM bandwidth unused).  With
# DW_OVERLAP_SMS = n > 0 (and in-place gradient accumulation, dp.FlatGradients(fused_accumulate=True)) the GEMM +
# scatter launches go to a side stream, and every backward chain launched while one of them is in flight leaves n SMs
# free (mcf_chain_params_t.max_ctas).  ``join_weight_gradients`` makes the current stream wait for the side stream; it
# is called by FlatGradients before the all-reduce and by FusedAdam.step.  Works under CUDA-graph capture (fork / join
# of the captured stream).
DW_OVERLAP_SMS = int(__import__('os').environ.get('MCF_DW_OVERLAP_SMS', '32'))
_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}
_SIDE_PENDING: Dict[int, bool] = {}


def _side_stream(device) -> "torch.cuda.Stream":
    idx = torch.device(device).index or 0
    st = _SIDE_STREAMS.get(idx)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _SIDE_STREAMS[idx] = st
    return st


def overlap_active() -> bool:
    return DW_OVERLAP_SMS > 0 and ACCUMULATE_INTO_GRAD


def chain_cta_cap(device) -> int:
    """max_ctas of a backward chain launch: leave DW_OVERLAP_SMS SMs to the weight-gradient stream while it is busy."""
    idx = torch.device(device).index or 0
    if not overlap_active() or not _SIDE_PENDING.get(idx):
        return 0
    n_sm = torch.cuda.get_device_properties(device).multi_processor_count
    return max(2, n_sm - DW_OVERLAP_SMS) & ~1     # even: the NeRF chain runs on CTA pairs


def join_weight_gradients() -> None:
    """The current stream waits for every weight-gradient launch issued so far (no-op when nothing is pending)."""
    for idx, pending in list(_SIDE_PENDING.items()):
        if pending:
            with torch.cuda.device(idx):
                torch.cuda.current_stream().wait_stream(_SIDE_STREAMS[idx])
            _SIDE_PENDING[idx] = False


def rayfeat_images(rf: torch.Tensor, n_rows: int, rows_per_ray: int) -> torch.Tensor:
    """bf16 tile images [n_tiles][128][64] of the per-ray feature columns: the Q operand of the weight gradients of the
    folded input columns.  Depends only on (features, ray layout): built once per render call per layout and shared by
    all evaluations (coarse/fine pass x 5 flow evaluations + NeRF used to write one copy each)."""
    from .mlp import memoised, _tkey
    def make():
        out = torch.empty(_n_tiles(n_rows) * L.BLOCK_BYTES, dtype=torch.uint8, device=rf.device)
        L.check(L.lib().mcf_rayfeat_image(L.ptr(rf), C.c_int(rf.stride(0)), C.c_int(rf.shape[1]), C.c_longlong(n_rows),
                                          C.c_int(rows_per_ray), L.ptr(out), L.stream_ptr()), "mcf_rayfeat_image")
        return out
    return memoised(("rfimg", _tkey(rf), n_rows, rows_per_ray), (rf,), make)


def _run_grad_plan(model, gp: P.GradPlan, st, need_dx: bool, fwd_save: torch.Tensor, fwd_tile_bytes: int,
                   bwd_save: torch.Tensor, bwd_tile_bytes: int, n_tiles: int, d_head: torch.Tensor, wanted: set,
                   aux: Optional[torch.Tensor] = None):
    """Runs the weight-gradient GEMMs; returns the flat gradient buffer, or None when the gradients were
    accumulated in place."""
    dev = fwd_save.device
    if overlap_active():
        side, main = _side_stream(dev), torch.cuda.current_stream()
        side.wait_stream(main)
        for t in (fwd_save, bwd_save, d_head, aux):
            if t is not None:
                t.record_stream(side)
        with torch.cuda.stream(side):
            out = _run_grad_plan_here(model, gp, st, need_dx, fwd_save, fwd_tile_bytes, bwd_save, bwd_tile_bytes, n_tiles,
                                      d_head, wanted, aux)
        _SIDE_PENDING[torch.device(dev).index or 0] = True
        return out
    return _run_grad_plan_here(model, gp, st, need_dx, fwd_save, fwd_tile_bytes, bwd_save, bwd_tile_bytes, n_tiles, d_head,
                               wanted, aux)


def _run_grad_plan_here(model, gp: P.GradPlan, st, need_dx: bool, fwd_save: torch.Tensor, fwd_tile_bytes: int,
                        bwd_save: torch.Tensor, bwd_tile_bytes: int, n_tiles: int, d_head: torch.Tensor, wanted: set,
                        aux: Optional[torch.Tensor] = None):
    dev = fwd_save.device
    staging = torch.zeros(gp.staging_floats, device=dev)
    key = (need_dx, frozenset(wanted))
    jobs_dev = st.jobs_dev.get(key)
    if jobs_dev is None:
        jobs_dev = _dev_table(P.job_table(gp, wanted), dev)
        st.jobs_dev[key] = jobs_dev
    work = sum(2.0 * (j.n_i + j.q_cols) * n_tiles * L.TILE_ROWS + 4.0 * j.n_i * j.n_j for j in gp.jobs
               if any(n in wanted for n in j.params))
    with L.timed("dw_gemm", work, "byte"):
        L.check(L.lib().mcf_dw_gemm_batch(L.ptr(jobs_dev), C.c_int(len(gp.jobs)), L.ptr(fwd_save),
                                          C.c_longlong(fwd_tile_bytes), L.ptr(bwd_save), C.c_longlong(bwd_tile_bytes),
                                          L.ptr(aux), C.c_longlong(L.BLOCK_BYTES),
                                          L.ptr(staging), C.c_longlong(n_tiles), C.c_int(dw_ctas_per_job(staging.device)), L.stream_ptr()),
                "mcf_dw_gemm_batch")
    ncols, stride, hc = gp.head_colsum
    L.check(L.lib().mcf_colsum(L.ptr(d_head), C.c_longlong(d_head.shape[0]), C.c_int(stride), C.c_int(ncols),
                               C.c_void_p(staging.data_ptr() + 4 * hc), L.stream_ptr()), "mcf_colsum")
    unpack_dev = st.unpack_dev[need_dx]
    if ACCUMULATE_INTO_GRAD:
        params = dict(model.named_parameters())
        ptrs = (C.c_void_p * len(gp.unpack_targets))()
        for i, (name, inner) in enumerate(gp.unpack_targets):
            if name not in wanted:
                ptrs[i] = None
                continue
            prm = params[name]
            if prm.grad is None:
                prm.grad = torch.zeros_like(prm)
            if not prm.grad.is_contiguous():
                raise RuntimeError("in-place gradient accumulation needs contiguous .grad tensors")
            ptrs[i] = prm.grad.data_ptr() + 4 * inner
        L.check(L.lib().mcf_unpack_accumulate(L.ptr(unpack_dev), C.c_int(len(gp.unpack)), L.ptr(staging), ptrs,
                                              L.stream_ptr()), "mcf_unpack_accumulate")
        return None
    grads = torch.zeros(gp.total_floats, device=dev)
    L.check(L.lib().mcf_unpack(L.ptr(unpack_dev), C.c_int(len(gp.unpack)), L.ptr(staging), L.ptr(grads),
                               L.stream_ptr()), "mcf_unpack")
    return grads


def _param_grads(gp: P.GradPlan, flat: torch.Tensor, names: List[str], needs: List[bool]):
    out = []
    for n, need in zip(names, needs):
        if not need:
            out.append(None)
            continue
        shp = gp.param_shapes[n]
        off = gp.param_offsets[n]
        out.append(flat[off:off + int(np.prod(shp))].view(shp))
    return out


# ------------------------------------------------------------------------------------------------
# NeRF
# ------------------------------------------------------------------------------------------------
class _NeRFFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, pe, rows_per_ray, xyz, dense, ray_feat, names, *params):
        from .autograd_mlp import _rows
        M = _rows(xyz, dense)
        R = M // rows_per_ray
        dev = (dense if dense is not None else xyz).device
        pp = model._plan(False, True)
        plan = pp.plan
        cp = ops.chain_params(pp, M, rows_per_ray, R)
        xyz_c = None if xyz is None else xyz.detach().contiguous()
        keep = setup_input(cp, xyz_c, pe, None if dense is None else dense.detach(), model.in_channels_xyz)
        E = model._extra_dim()
        rf = None
        if E > 0:
            lin = model.extra_encoding[0]
            rf = ray_feat[:, :E].detach()
            if rf.stride(-1) != 1:
                rf = rf.contiguous()
            rb = fold_bias(lin.weight, model.W, lin.bias, rf)
            cp.raybias[0] = rb.data_ptr()
            keep += [rb, rf]
        nt = _n_tiles(M)
        save = torch.empty(nt * plan.save_tile_bytes, dtype=torch.uint8, device=dev)
        masks = torch.empty(nt * plan.mask_tile_words, dtype=torch.int32, device=dev)
        out = torch.empty(M, 4, device=dev)
        cp.out, cp.out_stride, cp.sigma_col = out.data_ptr(), 4, 3
        cp.save, cp.save_tile_bytes = save.data_ptr(), plan.save_tile_bytes
        cp.masks, cp.mask_tile_words = masks.data_ptr(), plan.mask_tile_words
        cp.x0_save_off = plan.offsets["save_x0"]
        ops.launch_chain(cp, "nerf_fwd", ops.linear_flops(model))
        ctx.model, ctx.pe, ctx.names, ctx.M, ctx.S = model, pe, names, M, rows_per_ray
        ctx.dense_mode = dense is not None
        ctx.need_dx = xyz is not None and xyz.requires_grad
        # reference call convention NeRF.forward(inputs) with inputs that carry grad (trainer_moco_flow.py:146-158)
        ctx.need_ddense = dense is not None and dense.requires_grad
        ctx.dense_shape = tuple(dense.shape) if dense is not None else None
        ctx.aux = rayfeat_images(rf, M, rows_per_ray) if rf is not None else None
        ctx.save_for_backward(save, masks, out)
        return out

    @staticmethod
    def backward(ctx, g_out):
        model, M = ctx.model, ctx.M
        save, masks, out = ctx.saved_tensors
        dev = out.device
        st = _train_state(model)
        fwd_plan = model._plan(False, True).plan
        need_dx = bool(ctx.need_dx) or bool(ctx.need_ddense)
        if need_dx not in st.bwd:
            bplan = P.nerf_backward_plan(model.D, model.W, model.in_channels_xyz, tuple(model.skips),
                                         model._extra_dim(), need_dx, fwd_plan)
            st.bwd[need_dx] = ops.PackedPlan(bplan, dev)
            shapes = {k: tuple(v.shape) for k, v in model.named_parameters()}
            st.grad[need_dx] = P.nerf_grad_plan(model.D, model.W, model.in_channels_xyz, tuple(model.skips),
                                                model._extra_dim(), shapes, fwd_plan, bplan)
            st.unpack_dev[need_dx] = _dev_table(st.grad[need_dx].unpack, dev)
        bpp = st.bwd[need_dx]
        bpp.repack(model._param_dict())
        bplan = bpp.plan
        nt = _n_tiles(M)
        cp = ops.chain_params(bpp, M, ctx.S, M // ctx.S)
        cp.max_ctas = chain_cta_cap(dev)
        cp.prologue = L.PRO_B_NERF
        g_out = g_out.contiguous()
        cp.g_out, cp.fwd_out = g_out.data_ptr(), out.data_ptr()
        cp.fwd_save, cp.fwd_save_tile_bytes = save.data_ptr(), fwd_plan.save_tile_bytes
        cp.fwd_masks, cp.fwd_mask_tile_words = masks.data_ptr(), fwd_plan.mask_tile_words
        cp.fwd_x0_off = fwd_plan.offsets["save_x0"]
        bsave = torch.empty(nt * bplan.save_tile_bytes, dtype=torch.uint8, device=dev)
        cp.save, cp.save_tile_bytes = bsave.data_ptr(), bplan.save_tile_bytes
        cp.x0_save_off = bplan.offsets["save_dye"]
        cp.dhead_save_off = bplan.offsets["save_dhead"]
        d_head = torch.empty(M, 4, device=dev)
        cp.d_head = d_head.data_ptr()
        d_xyz = d_dense = None
        if ctx.need_ddense:
            d_dense = torch.zeros(ctx.dense_shape, device=dev)
            cp.d_dense, cp.d_dense_stride, cp.dense_cols = d_dense.data_ptr(), d_dense.stride(0), model.in_channels_xyz
        elif need_dx:
            d_xyz = torch.empty(M, 3, device=dev)
            cp.d_xyz = d_xyz.data_ptr()
            ops.set_pe(cp, ctx.pe, model.in_channels_xyz, dev)
        first = 2.0 * model.in_channels_xyz * model.W * (1 + len([s for s in model.skips if s > 0]))
        ops.launch_chain(cp, "nerf_bwd_dx", ops.linear_flops(model) - (0.0 if need_dx else first))
        E = model._extra_dim()
        if d_dense is not None and E > 0 and ctx.dense_shape[1] > model.in_channels_xyz:
            # extra-feature columns of the input rows: d = dYe @ W_e[:, W:W+E], dYe read back from its saved bf16 image
            dye = ops.image_rows(bsave, nt, bplan.save_tile_bytes, bplan.offsets["save_dye"], model.W // 2)[:M]
            w_e = model.extra_encoding[0].weight.detach()[:, model.W:model.W + E]
            ncol = min(E, ctx.dense_shape[1] - model.in_channels_xyz)
            d_dense[:, model.in_channels_xyz:model.in_channels_xyz + ncol] = (dye @ w_e)[:, :ncol]
        needs = list(ctx.needs_input_grad[7:])
        wanted = {n for n, need in zip(ctx.names, needs) if need}
        pgrads = [None] * len(ctx.names)
        if wanted:
            gp = st.grad[need_dx]
            flat = _run_grad_plan(model, gp, st, need_dx, save, fwd_plan.save_tile_bytes, bsave,
                                  bplan.save_tile_bytes, nt, d_head, wanted, ctx.aux)
            if flat is not None:
                pgrads = _param_grads(gp, flat, ctx.names, needs)
        return (None, None, None, d_xyz, d_dense, None, None, *pgrads)


def nerf_autograd(model, xyz, pe, dense, ray_feat, rows_per_ray, sigma_only):
    names = [n for n, _ in model.named_parameters()]
    params = [p for _, p in model.named_parameters()]
    if sigma_only:
        # Differentiable sigma-only evaluation (the auxiliary density loss on point batches,
        # trainer/trainer_moco_flow.py:146-158,349-361): sigma does not depend on the extra feature, so this runs the
        # full differentiable program with a zero feature row and returns its sigma column; the colour branch then
        # receives exactly-zero gradients.
        M = int(dense.shape[0] if dense is not None else xyz.shape[0])
        E = model._extra_dim()
        dev = (dense if dense is not None else xyz).device
        zero_feat = torch.zeros(M // rows_per_ray, E, device=dev) if E > 0 else None
        out = _NeRFFn.apply(model, pe, rows_per_ray, xyz, dense, zero_feat, names, *params)
        return out[:, 3:4]
    return _NeRFFn.apply(model, pe, rows_per_ray, xyz, dense, ray_feat, names, *params)


# ------------------------------------------------------------------------------------------------
# NoF
# ------------------------------------------------------------------------------------------------
class _NoFFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, pe, rows_per_ray, xyz, dense, ray_feat, names, *params):
        M = int(xyz.shape[0])
        R = M // rows_per_ray
        dev = xyz.device
        pp = model._plan(True)
        plan = pp.plan
        cp = ops.chain_params(pp, M, rows_per_ray, R)
        xyz_c = xyz.detach().contiguous()
        keep = setup_input(cp, xyz_c, pe, None if dense is None else dense.detach(), model.in_channels_xyz)
        E, cx = model.extra_feat_dim, model.in_channels_xyz
        if E > 0:
            from .autograd_mlp import check_nof_ray_feat
            check_nof_ray_feat(model, ray_feat, R)
            rf = ray_feat[:, :E].detach()
            if rf.stride(-1) != 1:
                rf = rf.contiguous()
            for k, i in enumerate(P.folded_layers(model.D, tuple(model.skips))):
                lin = getattr(model, f"nof_encoding_{i+1}")[0]
                rb = fold_bias(lin.weight, cx, lin.bias, rf)
                cp.raybias[k] = rb.data_ptr()
                keep.append(rb)
            keep.append(rf)
        else:   # the layer-1 weight GEMM still reads a (zero) feature block
            rf = torch.zeros(R, 1, device=dev)
        nt = _n_tiles(M)
        save = torch.empty(nt * plan.save_tile_bytes, dtype=torch.uint8, device=dev)
        masks = torch.empty(nt * plan.mask_tile_words, dtype=torch.int32, device=dev)
        out = torch.empty(M, 3, device=dev)
        head_save = torch.empty(M, 12, device=dev)
        cp.out, cp.out_stride, cp.use_quat = out.data_ptr(), 3, int(model.use_quat)
        cp.head_save = head_save.data_ptr()
        cp.save, cp.save_tile_bytes = save.data_ptr(), plan.save_tile_bytes
        cp.masks, cp.mask_tile_words = masks.data_ptr(), plan.mask_tile_words
        cp.x0_save_off = plan.offsets["save_x0"]
        ops.launch_chain(cp, "nof_fwd", ops.linear_flops(model))
        ctx.model, ctx.pe, ctx.names, ctx.M, ctx.S = model, pe, names, M, rows_per_ray
        ctx.aux = rayfeat_images(rf, M, rows_per_ray)
        ctx.need_dx = xyz.requires_grad
        ctx.save_for_backward(save, masks, head_save)
        return out

    @staticmethod
    def backward(ctx, g_out):
        model, M = ctx.model, ctx.M
        save, masks, head_save = ctx.saved_tensors
        dev = save.device
        st = _train_state(model)
        fwd_plan = model._plan(True).plan
        need_dx = bool(ctx.need_dx)
        if need_dx not in st.bwd:
            bplan = P.nof_backward_plan(model.D, model.W, model.in_channels_xyz, tuple(model.skips),
                                        model.extra_feat_dim, bool(model.use_quat), need_dx, fwd_plan)
            st.bwd[need_dx] = ops.PackedPlan(bplan, dev)
            shapes = {k: tuple(v.shape) for k, v in model.named_parameters()}
            st.grad[need_dx] = P.nof_grad_plan(model.D, model.W, model.in_channels_xyz, tuple(model.skips),
                                               model.extra_feat_dim, bool(model.use_quat), shapes, fwd_plan, bplan)
            st.unpack_dev[need_dx] = _dev_table(st.grad[need_dx].unpack, dev)
        bpp = st.bwd[need_dx]
        bpp.repack(model._param_dict())
        bplan = bpp.plan
        nt = _n_tiles(M)
        cp = ops.chain_params(bpp, M, ctx.S, M // ctx.S)
        cp.max_ctas = chain_cta_cap(dev)
        cp.prologue = L.PRO_B_NOF
        cp.use_quat = int(model.use_quat)
        g_out = g_out.contiguous()
        cp.g_out, cp.head_save = g_out.data_ptr(), head_save.data_ptr()
        cp.fwd_save, cp.fwd_save_tile_bytes = save.data_ptr(), fwd_plan.save_tile_bytes
        cp.fwd_masks, cp.fwd_mask_tile_words = masks.data_ptr(), fwd_plan.mask_tile_words
        cp.fwd_x0_off = fwd_plan.offsets["save_x0"]
        bsave = torch.empty(nt * bplan.save_tile_bytes, dtype=torch.uint8, device=dev)
        cp.save, cp.save_tile_bytes = bsave.data_ptr(), bplan.save_tile_bytes
        cp.x0_save_off = bplan.offsets["save_ghead"]
        d_head = torch.zeros(M, 12, device=dev)
        cp.d_head = d_head.data_ptr()
        d_xyz = None
        if need_dx:
            d_xyz = torch.empty(M, 3, device=dev)
            cp.d_xyz = d_xyz.data_ptr()
            ops.set_pe(cp, ctx.pe, model.in_channels_xyz, dev)
        first = 2.0 * model.in_channels_xyz * model.W * (1 + len([s for s in model.skips if s > 0]))
        ops.launch_chain(cp, "nof_bwd_dx", ops.linear_flops(model) - (0.0 if need_dx else first))
        needs = list(ctx.needs_input_grad[7:])
        wanted = {n for n, need in zip(ctx.names, needs) if need}
        pgrads = [None] * len(ctx.names)
        if wanted:
            gp = st.grad[need_dx]
            flat = _run_grad_plan(model, gp, st, need_dx, save, fwd_plan.save_tile_bytes, bsave,
                                  bplan.save_tile_bytes, nt, d_head, wanted, ctx.aux)
            if flat is not None:
                pgrads = _param_grads(gp, flat, ctx.names, needs)
        return (None, None, None, d_xyz, None, None, None, *pgrads)


def nof_autograd(model, xyz, pe, dense, ray_feat, rows_per_ray):
    if dense is not None and dense.requires_grad:
        raise NotImplementedError("gradients w.r.t. pre-embedded inputs are not provided; pass xyz + Embedding")
    if dense is not None and xyz.requires_grad:
        raise NotImplementedError("input-point gradients need the fused xyz encoder (pass pe=..., not dense=...)")
    names = [n for n, _ in model.named_parameters()]
    params = [p for _, p in model.named_parameters()]
    return _NoFFn.apply(model, pe, rows_per_ray, xyz, dense, ray_feat, names, *params)
