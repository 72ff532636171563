"""Host-side operator layer: torch tensors in, C-ABI kernel launches, torch tensors out.

PyTorch is used here for device memory, streams and autograd bookkeeping only; every arithmetic
step of the path runs in the CUDA library.  All functions raise if the inputs are not CUDA fp32
tensors -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os as _os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import plans as P


# MCF_DEBUG=1: every render_rays call ends with a device synchronisation and a check of the kernels' error flag
# (unknown epilogue / prologue for the launched kernel family, an expired bounded barrier wait, misaligned shared
# memory).  Without it the flag is checked wherever the host synchronises anyway: ``check_device()``.
DEBUG_SYNC = _os.environ.get("MCF_DEBUG", "0") == "1"


def check_device() -> None:
    """Synchronises and raises if any kernel of the library reported a device-side error since the last check."""
    flag = L.device_error_flag()
    if flag:
        raise L.MocoFlowLibraryError(f"device error flag {flag:#010x} (0xBADExxxx unknown epilogue, 0xBADFxxxx unknown "
                                     f"prologue, 0xDEADxxxx barrier wait expired, 0xA11Cxxxx shared-memory alignment)")


def _need_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("moco_flow_b200 operators run on CUDA tensors only (no CPU fallback)")
        if t.dtype != torch.float32:
            raise TypeError(f"expected float32 tensor, got {t.dtype}")


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.contiguous()


# ------------------------------------------------------------------------------------------------
# sampling
# ------------------------------------------------------------------------------------------------
def coarse_samples(rays: torch.Tensor, n_samples: int, perturb: float = 0.0,
                   perturb_rand: Optional[torch.Tensor] = None, use_disp: bool = False,
                   want_xyz: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """models/rendering.py:245-263."""
    rays = _c(rays)
    _need_cuda(rays, perturb_rand)
    R = rays.shape[0]
    t_steps = torch.linspace(0, 1, n_samples, device=rays.device)
    z = torch.empty(R, n_samples, device=rays.device)
    xyz = torch.empty(R, n_samples, 3, device=rays.device) if want_xyz else None
    L.check(L.lib().mcf_coarse_samples(L.ptr(rays), C.c_int(rays.shape[1]), L.ptr(t_steps), L.ptr(_c(perturb_rand)),
                                       C.c_float(perturb), C.c_int(int(use_disp)), C.c_int(R), C.c_int(n_samples),
                                       L.ptr(z), L.ptr(xyz), L.stream_ptr()), "mcf_coarse_samples")
    return z, xyz


def ray_points(rays: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """models/rendering.py:329-330."""
    rays, z = _c(rays), _c(z)
    _need_cuda(rays, z)
    R, S = z.shape
    xyz = torch.empty(R, S, 3, device=rays.device)
    L.check(L.lib().mcf_ray_points(L.ptr(rays), C.c_int(rays.shape[1]), L.ptr(z), C.c_int(R), C.c_int(S), L.ptr(xyz),
                                   L.stream_ptr()), "mcf_ray_points")
    return xyz


# ------------------------------------------------------------------------------------------------
# positional encoding (standalone)
# ------------------------------------------------------------------------------------------------
class _PEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, freqs, weights, table):
        x = _c(x)
        _need_cuda(x)
        B, Cin = x.shape
        nf = len(freqs)
        if nf > L.MAX_FREQS:
            raise ValueError(f"at most {L.MAX_FREQS} frequencies supported")
        out = torch.empty(B, Cin * (2 * nf + 1), device=x.device)
        L.check(L.lib().mcf_pe_fwd(L.ptr(x), C.c_longlong(B), C.c_int(Cin), C.c_int(nf), L.f32_array(freqs),
                                   L.f32_array(weights), L.ptr(table), L.ptr(out), C.c_int(out.shape[1]),
                                   L.stream_ptr()), "mcf_pe_fwd")
        ctx.save_for_backward(x)
        ctx.fw = (list(freqs), list(weights), table)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        freqs, weights, table = ctx.fw
        dy = _c(dy)
        dx = torch.empty_like(x)
        L.check(L.lib().mcf_pe_bwd(L.ptr(x), L.ptr(dy), C.c_longlong(x.shape[0]), C.c_int(x.shape[1]),
                                   C.c_int(len(freqs)), L.f32_array(freqs), L.f32_array(weights), L.ptr(table),
                                   C.c_int(dy.shape[1]), L.ptr(dx), L.stream_ptr()), "mcf_pe_bwd")
        return dx, None, None, None


def pe_forward(x: torch.Tensor, freqs: Sequence[float], weights: Sequence[float],
               table: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``table``: device {freq[], weight[]} copy (Embedding.device_table) that the kernel reads instead of the
    by-value arrays."""
    return _PEFn.apply(x, tuple(freqs), tuple(weights), table)


def ray_bias(weight: torch.Tensor, col_off: int, bias: Optional[torch.Tensor], feat: torch.Tensor) -> torch.Tensor:
    """out[r] = bias + weight[:, col_off:col_off+E] @ feat[r]  (fp32)."""
    weight, feat, bias = _c(weight.detach()), _c(feat.detach()), _c(None if bias is None else bias.detach())
    _need_cuda(weight, feat, bias)
    R, E = feat.shape
    N = weight.shape[0]
    out = torch.empty(R, N, device=feat.device)
    L.check(L.lib().mcf_ray_bias(L.ptr(weight), C.c_int(weight.shape[1]), C.c_int(col_off), L.ptr(bias), L.ptr(feat),
                                 C.c_int(E), C.c_int(E), C.c_int(R), C.c_int(N), L.ptr(out), L.stream_ptr()),
            "mcf_ray_bias")
    return out


# ------------------------------------------------------------------------------------------------
# alpha compositing
# ------------------------------------------------------------------------------------------------
class _CompositeFn(torch.autograd.Function):
    """models/rendering.py:158-190 with its analytic backward."""

    @staticmethod
    def forward(ctx, raw, z, dirs, noise, noise_std, background, act):
        # raw: (R,S,4) [rgb, sigma] or (R,S) sigma only
        raw, z, dirs, noise, background = _c(raw), _c(z), _c(dirs), _c(noise), _c(background)
        _need_cuda(raw, z, dirs, noise, background)
        R, S = z.shape
        full = raw.dim() == 3
        dev = z.device
        weights = torch.empty(R, S, device=dev)
        alphas = torch.empty(R, S, device=dev)
        opacity = torch.empty(R, device=dev)
        rgb = torch.empty(R, 3, device=dev) if full else None
        depth = torch.empty(R, device=dev) if full else None
        base = raw.data_ptr()
        sig_ptr = C.c_void_p(base + 12) if full else C.c_void_p(base)
        with L.timed("composite_fwd", R * (28.0 * S + 44.0) if full else R * (16.0 * S + 16.0), "byte"):
            L.check(L.lib().mcf_composite_fwd(
                sig_ptr, C.c_int(4 if full else 1), C.c_void_p(base if full else 0), C.c_int(4), L.ptr(z),
                L.ptr(dirs), C.c_int(dirs.shape[1]), L.ptr(noise), C.c_float(noise_std), L.ptr(background),
                C.c_int(act), C.c_int(R), C.c_int(S), L.ptr(weights), L.ptr(alphas), L.ptr(rgb), L.ptr(depth),
                L.ptr(opacity), L.stream_ptr()), "mcf_composite_fwd")
        ctx.save_for_backward(raw, z, dirs, noise, background)
        ctx.cfg = (noise_std, act, full)
        ctx.mark_non_differentiable(alphas)
        if full:
            return rgb, depth, weights, alphas, opacity
        return weights, alphas, opacity

    @staticmethod
    def backward(ctx, *grads):
        raw, z, dirs, noise, background = ctx.saved_tensors
        noise_std, act, full = ctx.cfg
        if full:
            g_rgb, g_depth, g_w, _, g_op = grads
        else:
            g_w, _, g_op = grads
            g_rgb = g_depth = None
        R, S = z.shape
        d_raw = torch.empty_like(raw)
        base, dbase = raw.data_ptr(), d_raw.data_ptr()
        with L.timed("composite_bwd", R * (40.0 * S + 28.0), "byte"):
            L.check(L.lib().mcf_composite_bwd(
                C.c_void_p(base + 12 if full else base), C.c_int(4 if full else 1), C.c_void_p(base if full else 0),
                C.c_int(4), L.ptr(z), L.ptr(dirs), C.c_int(dirs.shape[1]), L.ptr(noise), C.c_float(noise_std),
                L.ptr(background), C.c_int(act), C.c_int(R), C.c_int(S), L.ptr(_c(g_rgb)), L.ptr(_c(g_depth)),
                L.ptr(_c(g_op)), L.ptr(_c(g_w)), C.c_void_p(dbase + 12 if full else dbase), C.c_int(4 if full else 1),
                C.c_void_p(dbase if full else 0), C.c_int(4), L.stream_ptr()), "mcf_composite_bwd")
        return d_raw, None, None, None, None, None, None


def composite(raw, z, dirs, noise, noise_std, background, activate_type):
    if activate_type not in L.ACT:
        raise ValueError("activation layer type: %s not support" % activate_type)  # rendering.py:174
    return _CompositeFn.apply(raw, z, dirs, noise, float(noise_std), background, L.ACT[activate_type])


# ------------------------------------------------------------------------------------------------
# sample_pdf (+ sort-merge)
# ------------------------------------------------------------------------------------------------
# How the cdf of sample_pdf is built when the caller passes weights (exactness contract):
#   False (default) -- "reference order": the reference's own four torch ops (rendering.py:20-23) on the same device,
#                      so the cdf -- and with it every index -- is bit-identical to the reference run on this GPU;
#                      the kernel fuses everything after the cdf (search, gather, lerp, sort-merge).
#   True            -- the kernel builds the cdf itself in a fixed order (fp64 total, fp64 running sum rounded per
#                      entry); one pass less, indices can differ from torch's where u is within 1 ulp of a cdf entry.
FUSED_CDF = _os.environ.get("MCF_FUSED_CDF", "0") == "1"


def reference_cdf(weights: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """rendering.py:20-23, literally (same ops, same order, same device)."""
    weights = weights + eps
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)


def sample_pdf_raw(bins: torch.Tensor, weights: Optional[torch.Tensor], u: torch.Tensor, eps: float = 1e-5,
                   bins_are_z: bool = False, w_offset: int = 0, n_bins: Optional[int] = None,
                   cdf: Optional[torch.Tensor] = None, z_coarse: Optional[torch.Tensor] = None,
                   want_samples: bool = True, want_inds: bool = False, want_cdf: bool = False,
                   fused_cdf: Optional[bool] = None):
    """models/rendering.py:5-46 (+:326).  ``weights`` may be a wider row of which columns
    [w_offset, w_offset+n_bins) are the bin weights (the reference passes weights[:, 1:-1])."""
    if n_bins is None:
        n_bins = (weights.shape[1] - w_offset) if weights is not None else cdf.shape[1] - 1
    if cdf is None and not (FUSED_CDF if fused_cdf is None else fused_cdf):
        _need_cuda(weights)
        cdf, weights = reference_cdf(weights[:, w_offset:w_offset + n_bins], eps), None
    bins, weights, u, cdf, z_coarse = _c(bins), _c(weights), _c(u), _c(cdf), _c(z_coarse)
    _need_cuda(bins, weights, u, cdf, z_coarse)
    R, n_imp = u.shape
    dev = u.device
    samples = torch.empty(R, n_imp, device=dev) if want_samples else None
    inds = torch.empty(R, n_imp, device=dev, dtype=torch.int32) if want_inds else None
    cdf_out = torch.empty(R, n_bins + 1, device=dev) if want_cdf else None
    merged = torch.empty(R, z_coarse.shape[1] + n_imp, device=dev) if z_coarse is not None else None
    wptr = C.c_void_p(0 if weights is None else weights.data_ptr() + 4 * w_offset)
    n_c = 0 if z_coarse is None else z_coarse.shape[1]
    with L.timed("sample_pdf", R * 4.0 * ((n_bins + 1) + n_bins + 2 * n_imp + 2 * n_c), "byte"):
        L.check(L.lib().mcf_sample_pdf(
            L.ptr(bins), C.c_int(bins.shape[1]), C.c_int(int(bins_are_z)), wptr,
            C.c_int(0 if weights is None else weights.shape[1]), L.ptr(cdf), C.c_int(0 if cdf is None else cdf.shape[1]),
            L.ptr(u), C.c_int(u.shape[1]), C.c_float(eps), C.c_int(R), C.c_int(n_bins), C.c_int(n_imp), L.ptr(z_coarse),
            C.c_int(0 if z_coarse is None else z_coarse.shape[1]), C.c_int(0 if z_coarse is None else z_coarse.shape[1]),
            L.ptr(samples), L.ptr(inds), L.ptr(cdf_out), L.ptr(merged), L.stream_ptr()), "mcf_sample_pdf")
    return samples, inds, cdf_out, merged


# ------------------------------------------------------------------------------------------------
# flow-consistency residual
# ------------------------------------------------------------------------------------------------
class _ResidualFn(torch.autograd.Function):
    """Per-sample mean_3|a-b| (models/rendering.py:310-311 before the mask); grad flows to b only."""

    @staticmethod
    def forward(ctx, a, b, alphas):
        a, b, alphas = _c(a), _c(b), _c(alphas)
        _need_cuda(a, b, alphas)
        M = alphas.numel()
        resid = torch.empty(alphas.shape, device=a.device)
        L.check(L.lib().mcf_masked_l1_fwd(L.ptr(a), L.ptr(b), L.ptr(alphas), C.c_float(0.01), C.c_longlong(M),
                                          L.ptr(resid), C.c_void_p(0), C.c_void_p(0), L.stream_ptr()),
                "mcf_masked_l1_fwd")
        ctx.save_for_backward(a, b, alphas)
        return resid

    @staticmethod
    def backward(ctx, g):
        a, b, alphas = ctx.saved_tensors
        g = _c(g)
        d_b = torch.empty_like(b)
        L.check(L.lib().mcf_masked_l1_bwd(L.ptr(a), L.ptr(b), L.ptr(alphas), C.c_float(0.01),
                                          C.c_longlong(alphas.numel()), L.ptr(g), C.c_void_p(0), C.c_void_p(0),
                                          C.c_float(1.0), L.ptr(d_b), L.stream_ptr()), "mcf_masked_l1_bwd")
        return None, d_b, None


class _MaskedMeanFn(torch.autograd.Function):
    """Sync-free fused form: mean of the residual over alphas>=0.01 (all samples if none), as a (1,)
    tensor -- torch.mean of it equals torch.mean of the reference's dynamic-length vector.  ``stats`` holds the
    {masked sum, masked count, total sum, total count} of the residual (already summed over the data-parallel ranks
    when there are several); ``grad_mul`` is the world size (the ranks' gradients are averaged afterwards)."""

    @staticmethod
    def forward(ctx, a, b, alphas, stats, grad_mul):
        out = torch.empty(1, device=a.device)
        L.check(L.lib().mcf_masked_l1_finalize(L.ptr(stats), L.ptr(out), L.stream_ptr()), "mcf_masked_l1_finalize")
        ctx.save_for_backward(a, b, alphas, stats)
        ctx.grad_mul = float(grad_mul)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b, alphas, stats = ctx.saved_tensors
        g = _c(g)
        d_b = torch.empty_like(b)
        L.check(L.lib().mcf_masked_l1_bwd(L.ptr(a), L.ptr(b), L.ptr(alphas), C.c_float(0.01),
                                          C.c_longlong(alphas.numel()), C.c_void_p(0), L.ptr(g), L.ptr(stats),
                                          C.c_float(ctx.grad_mul), L.ptr(d_b), L.stream_ptr()), "mcf_masked_l1_bwd")
        return None, d_b, None, None, None


# Data parallelism: (process group, world size) over which the masked residual statistics are summed before the mean
# is taken, so that every rank computes the mean over the whole ray batch of the step (set by
# dp.enable_global_residual_means; None = this process's rays only).
RESIDUAL_DP = None


class ResidualMeans:
    """The fused flow-residual means of one render_rays call: every residual's statistics go into one row of a
    [n][4] fp64 tensor, ONE all-reduce sums the rows over the data-parallel ranks, then each mean is finalised."""

    def __init__(self, device, capacity: int = 4):
        self.stats = torch.zeros(capacity, 4, dtype=torch.float64, device=device)
        self.pending = []

    def add(self, key: str, x_obs, x_rec, alphas) -> None:
        a, b, alphas = _c(x_obs.detach()), _c(x_rec), _c(alphas.detach())
        _need_cuda(a, b, alphas)
        row = self.stats[len(self.pending)]
        L.check(L.lib().mcf_masked_l1_fwd(L.ptr(a), L.ptr(b.detach()), L.ptr(alphas), C.c_float(0.01),
                                          C.c_longlong(alphas.numel()), C.c_void_p(0), L.ptr(row), C.c_void_p(0),
                                          L.stream_ptr()), "mcf_masked_l1_fwd")
        self.pending.append((key, a, b, alphas, row))

    def finish(self, result: dict) -> None:
        if not self.pending:
            return
        mul = 1.0
        if RESIDUAL_DP is not None:
            import torch.distributed as dist
            group, world = RESIDUAL_DP
            if world > 1:
                dist.all_reduce(self.stats, op=dist.ReduceOp.SUM, group=group)
                mul = float(world)
        for key, a, b, alphas, row in self.pending:
            result[key] = _MaskedMeanFn.apply(a, b, alphas, row, mul)
        self.pending = []


def flow_residual(x_obs, x_rec, alphas, fused_mean: bool = False):
    """models/rendering.py:306-311.  ``fused_mean=False`` reproduces the reference's dynamic-length
    vector (one host sync, like the reference's torch.any); ``True`` returns the (1,) masked mean."""
    x_obs = x_obs.detach()
    if fused_mean:
        rm = ResidualMeans(x_obs.device, 1)
        rm.add("r", x_obs, x_rec, alphas)
        out = {}
        rm.finish(out)
        return out["r"]
    resid = _ResidualFn.apply(x_obs, x_rec, alphas)
    mask = alphas >= 0.01
    if not bool(torch.any(mask)):
        mask = torch.ones_like(mask)
    return resid[mask]


# ------------------------------------------------------------------------------------------------
# fused MLP chains
# ------------------------------------------------------------------------------------------------
# Width-256 chains run on CTA pairs (tcgen05 cta_group::2, see mcf_chain_params_t.cta_pair): bit-identical to the
# one-CTA-per-tile-pair path (tests/test_gpu_chain.py::test_nerf_cta_pair_matches_single), 8-9 % faster on the training
# chains, equal on inference.  MCF_CTA_PAIR=0 selects the single-CTA path.
CTA_PAIR = int(_os.environ.get("MCF_CTA_PAIR", "1"))
# training chains: signal the next layer's MMA before issuing the bulk store of the saved operand image
EARLY_ARRIVE = int(_os.environ.get("MCF_EARLY_ARRIVE", "1"))


# bumped by writers that update parameters behind autograd's back (optim.FusedAdam writes through raw pointers, which
# does not advance the tensors' version counters): packed bf16 weight images are rebuilt when it changes
PARAM_EPOCH = 0


def bump_param_epoch() -> None:
    global PARAM_EPOCH
    PARAM_EPOCH += 1


class PackedPlan:
    """A plan's device-side tables plus the packed weight buffers of one module."""

    def __init__(self, plan: P.Plan, device):
        self.plan = plan
        self.device = device
        as_dev = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(device)
        self.pack_tab = as_dev(plan.pack)
        self.chunk_tab = as_dev(plan.chunks)
        self.round_tab = as_dev(plan.rounds)
        self.wpack = torch.zeros(max(plan.wpack_bytes, 16), dtype=torch.uint8, device=device)
        self.consts = torch.zeros(plan.n_consts, dtype=torch.float32, device=device)
        self.versions = None

    def repack(self, params: Dict[str, torch.Tensor]) -> None:
        tensors = [params[n] for n in self.plan.tensor_names]
        versions = (PARAM_EPOCH,) + tuple((t.data_ptr(), t._version) for t in tensors)
        if versions == self.versions:
            return
        for t in tensors:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("module parameters must be contiguous CUDA float32 tensors")
        arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        L.check(L.lib().mcf_pack(L.ptr(self.pack_tab), C.c_int(len(self.plan.pack)), arr, C.c_int(len(tensors)),
                                 L.ptr(self.wpack), L.ptr(self.consts), L.stream_ptr()), "mcf_pack")
        self.versions = versions


def chain_params(pp: PackedPlan, n_rows: int, rows_per_ray: int, n_rays: int) -> L.ChainParams:
    cp = L.ChainParams()
    cp.chunks, cp.rounds = pp.chunk_tab.data_ptr(), pp.round_tab.data_ptr()
    cp.n_chunks, cp.n_rounds = len(pp.plan.chunks), len(pp.plan.rounds)
    cp.width = pp.plan.width
    cp.wpack, cp.consts = pp.wpack.data_ptr(), pp.consts.data_ptr()
    cp.n_rows, cp.rows_per_ray, cp.n_rays = n_rows, rows_per_ray, n_rays
    cp.x0_save_off = L.NONE
    cp.fwd_x0_off = cp.fwd_he_off = L.NONE
    cp.extra_save_off = cp.dhead_save_off = L.NONE
    cp.out_stride, cp.sigma_col = 4, 3
    cp.cta_pair = CTA_PAIR if pp.plan.width == 256 else 0
    cp.program_kind = pp.plan.kind
    cp.wpack_bytes, cp.resident = pp.plan.wpack_bytes, int(pp.plan.resident)
    cp.reserved0 = EARLY_ARRIVE
    return cp


def set_pe(cp: L.ChainParams, pe, pad_to: int, device) -> None:
    """Encoder tables of a chain launch: by value (plain C-ABI callers) and as the Embedding's device table, which
    is what the kernel reads (stays current under CUDA-graph replay, see Embedding.device_table)."""
    freqs, weights = pe.frequencies(), pe.multipliers()
    if len(freqs) > 10:
        raise ValueError("the fused xyz encoder supports at most 10 frequencies (63 channels)")
    cp.pe_n_freqs, cp.pe_pad_to = len(freqs), pad_to
    for i, (f, w) in enumerate(zip(freqs, weights)):
        cp.pe_freq[i], cp.pe_weight[i] = float(f), float(w)
    cp.pe_table = pe.device_table(device).data_ptr()


TIMING = None  # when a dict: tag -> list of [n_ctas][16] in-kernel cycle-counter tensors (scripts/chain_timing.py)


def launch_chain(cp: L.ChainParams, tag: str = "chain", flops_per_row: float = 0.0) -> None:
    buf = None
    if TIMING is not None:
        buf = torch.zeros(256, 16, dtype=torch.int64, device="cuda")
        cp.timing = buf.data_ptr()
    with L.timed(tag, flops_per_row * cp.n_rows, "flop"):
        L.check(L.lib().mcf_chain_launch(C.byref(cp), L.stream_ptr()), "mcf_chain_launch")
    if buf is not None:
        TIMING.setdefault(tag, []).append(buf)


def image_rows(buf: torch.Tensor, n_tiles: int, tile_bytes: int, off: int, n_cols: int) -> torch.Tensor:
    """A saved bf16 operand image ([n_tiles] records, blocks of [128 rows][64 cols], 16-byte chunk c of row r stored at
    chunk c ^ (r & 7)) back as a row-major fp32 (n_tiles*128, n_cols) tensor.  Plumbing for rarely used gradient paths
    (pre-embedded inputs); the hot paths consume the images on the tensor cores."""
    nb = (n_cols + 63) // 64
    rec = buf.view(n_tiles, tile_bytes)[:, off:off + nb * L.BLOCK_BYTES].contiguous()
    x = rec.view(torch.bfloat16).view(n_tiles, nb, 128, 8, 8)            # [tile][block][row][stored chunk][8]
    r = torch.arange(128, device=buf.device).view(128, 1)
    c = torch.arange(8, device=buf.device).view(1, 8)
    src = (c ^ (r & 7)).view(1, 1, 128, 8, 1).expand(n_tiles, nb, 128, 8, 8)
    x = torch.gather(x, 3, src)                                          # logical chunk c <- stored chunk c ^ (r & 7)
    return x.permute(0, 2, 1, 3, 4).reshape(n_tiles * 128, nb * 64)[:, :n_cols].float()


def linear_flops(module) -> float:
    """Algorithmic FLOPs per row of a module's Linear layers (2*in*out, un-padded reference shapes)."""
    return float(sum(2 * m.in_features * m.out_features for m in module.modules() if isinstance(m, torch.nn.Linear)))


def dw_gemm(p_base: torch.Tensor, p_tile_bytes: int, p_off: int, p_cols: int, q_base: torch.Tensor,
            q_tile_bytes: int, q_off: int, q_cols: int, out: torch.Tensor, n_i: int, n_j: int, n_tiles: int,
            colsum: Optional[torch.Tensor] = None, max_ctas: int = 0) -> None:
    dp = L.DwParams()
    dp.p_base, dp.p_tile_bytes, dp.p_off, dp.p_cols = p_base.data_ptr(), p_tile_bytes, p_off, p_cols
    dp.q_base, dp.q_tile_bytes, dp.q_off, dp.q_cols = q_base.data_ptr(), q_tile_bytes, q_off, q_cols
    dp.out, dp.ld_out, dp.n_i, dp.n_j = out.data_ptr(), out.stride(0), n_i, n_j
    dp.colsum_p = 0 if colsum is None else colsum.data_ptr()
    dp.n_tiles, dp.max_ctas = n_tiles, max_ctas
    # HBM-bound by construction: every P / Q image byte is read once (2*(in+out) B per sample)
    with L.timed("dw_gemm", 2.0 * (n_i + q_cols) * n_tiles * L.TILE_ROWS + 4.0 * n_i * n_j, "byte"):
        L.check(L.lib().mcf_dw_gemm(C.byref(dp), L.stream_ptr()), "mcf_dw_gemm")
