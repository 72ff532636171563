"""Layer programs ("plans") for the tcgen05 chain kernel.

A plan is three small tables (see include/moco_flow_b200.h):
  * pack   -- how to turn the fp32 nn.Linear tensors into bf16 128B-swizzled weight chunk images
              (one image = the B operand of up to four K=16 MMAs) and fp32 constants,
  * chunks -- the order in which the kernel streams those images and which accumulator columns
              and A-operand block each one feeds,
  * rounds -- groups of chunks followed by one epilogue (bias/activation/head).
Plans depend only on the module shapes, so they are built once per module and cached.

Shapes follow models/nerf.py:28-59 and models/nof.py:40-53 of the reference.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L

BLK = L.BLOCK_BYTES
FUSE_N256 = os.environ.get("MCF_FUSE_N256", "1") != "0"
# NoF programs (W = 128, <= 144 KB of packed weights) run on the resident-weight kernel: weights copied into shared
# memory once per CTA, no per-tile weight stream.  MCF_NOF_RESIDENT=0 builds the streamed (ring) programs instead.
NOF_RESIDENT = os.environ.get("MCF_NOF_RESIDENT", "1") != "0"
# which resident kernel: "ts" = nof_chain.cu (activations in tensor memory, mcf_chain_params_t.resident = 2),
# "smem" = chain.cu's resident variant (resident = 1)
NOF_KERNEL = os.environ.get("MCF_NOF_KERNEL", "ts")
# Width-256 programs: lay the weight tiles of a layer out half by half ([kb0 nh0][kb1 nh0]..[kb0 nh1][kb1 nh1]..), so
# that ONE CTA's share of consecutive k-blocks is contiguous in the packed stream; the CTA-pair kernel then fetches two
# k-blocks with one 32 KB bulk copy instead of two 16 KB ones (a copy issue costs its thread ~330 clk whatever the size:
# 24 B/clk with 16 KB copies, while a pair tile consumes 32 B/clk per CTA at the MMA's peak rate).
PAIR_MERGE = os.environ.get("MCF_PAIR_MERGE", "1") != "0"
RES_BYTES = 147456


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b


@dataclass
class Plan:
    width: int
    tensor_names: List[str]
    pack: np.ndarray
    chunks: np.ndarray
    rounds: np.ndarray
    wpack_bytes: int
    n_consts: int
    save_tile_bytes: int = 0
    mask_tile_words: int = 0
    offsets: Dict[str, int] = field(default_factory=dict)   # named save / mask / const offsets
    n_raybias: int = 0
    kind: int = 0          # 0: NeRF program, 1: NoF program (selects the kernel instantiation)
    resident: int = 0      # mcf_chain_params_t.resident: 0 streamed weights, 1 resident (smem activations), 2 TMEM-resident


class _Builder:
    def __init__(self, width: int):
        self.width = width
        self.pair_layout = PAIR_MERGE and width == 256
        self.names: List[str] = []
        self.pack: List[tuple] = []
        self.chunks: List[tuple] = []
        self.rounds: List[tuple] = []
        self.wbytes = 0
        self.nconst = 0
        self.save_bytes = 0
        self.mask_words = 0
        self.offsets: Dict[str, int] = {}

    def tensor(self, name: str) -> int:
        if name not in self.names:
            self.names.append(name)
        return self.names.index(name)

    def const(self, name: str, row0: int, nrows: int, col0: int, ncols: int, ld: int, pad_to: Optional[int] = None) -> int:
        """Copies src[row0:row0+nrows, col0:col0+ncols] (row-major, tightly) into consts; returns float offset."""
        n = nrows * ncols
        total = _ceil(max(n, pad_to or 0), 4) * 4
        off = self.nconst
        self.pack.append((off, total, 1, self.tensor(name), row0, nrows, col0, ncols, ld, 0))
        # kind 1 walks i -> (r = i / ncols, c = i % ncols); entries past nrows are zero-filled
        self.nconst += total
        return off

    def image(self, name: str, row0: int, nrows: int, col0: int, ncols: int, ld: int, transposed: bool,
              rows_padded: int) -> tuple:
        assert ncols <= 64 and nrows <= rows_padded and rows_padded % 8 == 0
        off, nbytes = self.wbytes, rows_padded * 128
        self.pack.append((off, nbytes, 0, self.tensor(name), row0, nrows, col0, ncols, ld, int(transposed)))
        self.wbytes += nbytes
        return off, nbytes

    def chunk(self, img: tuple, a_buf: int, a_kblock: int, ksteps: int, n: int, acc_col: int, init: bool) -> None:
        assert 1 <= ksteps <= 4 and n % 16 == 0 and 16 <= n <= 256
        self.chunks.append([img[0], img[1], a_buf, a_kblock, ksteps, 1 if init else 0, n, acc_col])
        # Two consecutive 128-row halves of the same [256 x 64] weight tile landing in an (even, odd) pair of ring
        # stages are one contiguous 32 KB K-major tile: mark the first so the kernel issues N=256 instructions
        # (halves the A-operand shared-memory reads per FLOP).
        # (In the ring kernels the two halves become adjacent in shared memory whatever their stream offsets; the
        # resident-weight kernel reads the stream itself and needs them adjacent there.)
        if FUSE_N256 and len(self.chunks) >= 2 and len(self.chunks) % 2 == 0:
            a, b = self.chunks[-2], self.chunks[-1]
            if (a[2], a[3], a[4], a[5] & 1, a[6]) == (b[2], b[3], b[4], b[5] & 1, b[6]) and a[6] == 128 \
                    and a[1] == BLK and b[1] == BLK and b[7] == a[7] + 128 \
                    and (b[0] == a[0] + BLK or self.pair_layout):
                a[5] |= 2

    def tiles(self, img_args, kbs: Sequence[int], NH: int, a_buf: int, acc0: int, init_kb: int) -> None:
        """The [NH*128 x 64] weight tiles of the k-blocks ``kbs`` of one layer: images in stream order, chunks in
        consumption order (kb major, then the 128-row halves).  ``img_args(kb, nh)`` -> arguments of ``image``."""
        imgs = {}
        if self.pair_layout and NH == 2:
            for nh in range(NH):
                for kb in kbs:
                    imgs[(kb, nh)] = self.image(*img_args(kb, nh))
        else:
            for kb in kbs:
                for nh in range(NH):
                    imgs[(kb, nh)] = self.image(*img_args(kb, nh))
        for kb in kbs:
            for nh in range(NH):
                self.chunk(imgs[(kb, nh)], a_buf, kb, 4, 128, acc0 + nh * 128, init=(kb == init_kb))

    def round(self, epi: int, n_out: int, acc_col: int, chunk_begin: int, raybias: int = -1, const_off: int = 0,
              aux_off: int = 0, save_off: int = L.NONE, mask_off: int = L.NONE) -> None:
        # flag bit 3 on the first chunk of a pair tile: this CTA's halves of this k-block and of the next one are
        # adjacent in the stream (pair layout) -> the CTA-pair kernel may fetch both with one 32 KB copy
        ch = self.chunks
        for c in range(chunk_begin, len(ch) - 3):
            a, a1, b, b1 = ch[c], ch[c + 1], ch[c + 2], ch[c + 3]
            if (a[5] & 2) and (b[5] & 2) and (c - chunk_begin) % 2 == 0 and a[2] == b[2] and b[3] == a[3] + 1 \
                    and a[4] == 4 and b[4] == 4 and not (b[5] & 1) and b[0] == a[0] + BLK and b1[0] == a1[0] + BLK:
                a[5] |= 8
        self.rounds.append((epi, n_out, acc_col, chunk_begin, len(self.chunks), raybias, const_off, aux_off,
                            save_off, mask_off, 0))

    def save_slot(self, key: str, n_blocks: int) -> int:
        off = self.save_bytes
        self.offsets["save_" + key] = off
        self.save_bytes += n_blocks * BLK
        return off

    def mask_slot(self, key: str, n_cols: int) -> int:
        off = self.mask_words
        self.offsets["mask_" + key] = off
        self.mask_words += _ceil(n_cols, 32) * 128
        return off

    def finish(self, n_raybias: int = 0, kind: int = 0, resident: int = 0) -> Plan:
        self.chunks = [tuple(c) for c in self.chunks]
        if len(self.chunks) > 128 or len(self.rounds) > 24 or len(self.names) > 32:
            raise ValueError(f"network too deep for one fused launch: {len(self.chunks)} weight chunks (max 128), "
                             f"{len(self.rounds)} layer rounds (max 24), {len(self.names)} parameter tensors (max 32)")
        assert not resident or self.wbytes <= RES_BYTES
        return Plan(self.width, list(self.names),
                    np.array(self.pack, dtype=L.PACK_DT), np.array(self.chunks, dtype=L.CHUNK_DT),
                    np.array(self.rounds, dtype=L.ROUND_DT), self.wbytes, max(self.nconst, 4),
                    self.save_bytes, self.mask_words, dict(self.offsets), n_raybias, kind, resident)


def _check_common(W: int, cx: int) -> None:
    if W not in (128, 256):
        raise ValueError(f"fused MLP kernels support hidden width 128 or 256, got {W}")
    if cx > 64:
        raise ValueError(f"fused MLP kernels support in_channels_xyz <= 64, got {cx}")


def _trunk_sources(i: int, skips: Sequence[int], cx: int, nkb: int, skip_extra: int):
    """[(a_buf, a_kblock, ksteps, weight col0, ncols)] of trunk layer i (0-based)."""
    kx = _ceil(cx, 16)
    if i == 0:
        return [(0, 0, kx, 0, cx)]
    src = []
    base = 0
    if i in skips:
        src.append((0, 0, kx, 0, cx))
        base = cx + skip_extra
    src += [(1, kb, 4, base + 64 * kb, 64) for kb in range(nkb)]
    return src


def nerf_forward_plan(D: int, W: int, cx: int, skips: Sequence[int], extra_dim: int, sigma_only: bool,
                      training: bool) -> Plan:
    """models/nerf.py:61-102 as a chain program."""
    _check_common(W, cx)
    b = _Builder(W)
    nkb, NH = W // 64, W // 128
    if training:
        b.save_slot("x0", 1)
    for i in range(D):
        wname, bname = f"xyz_encoding_{i+1}.0.weight", f"xyz_encoding_{i+1}.0.bias"
        ld = cx if i == 0 else (W + cx if i in skips else W)
        c0 = len(b.chunks)
        kx = _ceil(cx, 16)
        has_x0 = i == 0 or i in skips
        if has_x0:
            for nh in range(NH):
                b.chunk(b.image(wname, nh * 128, 128, 0, cx, ld, False, 128), 0, 0, kx, 128, nh * 128, init=True)
        if i > 0:
            base = cx if i in skips else 0
            b.tiles(lambda kb, nh: (wname, nh * 128, 128, base + 64 * kb, 64, ld, False, 128), list(range(nkb)), NH, 1, 0,
                    -1 if has_x0 else 0)
        boff = b.const(bname, 0, 1, 0, W, W)
        last = i == D - 1
        aux = 0
        if last:
            aux = b.const("sigma.weight", 0, 1, 0, W, W)
            b.const("sigma.bias", 0, 1, 0, 1, 1)  # lands at aux + W
        save = b.save_slot(f"h{i+1}", nkb) if training else L.NONE
        mask = b.mask_slot(f"h{i+1}", W) if training else L.NONE
        b.round(L.EPI_RELU_SIGMA if last else L.EPI_RELU, W, 0, c0, const_off=boff, aux_off=aux, save_off=save,
                mask_off=mask)
    if not sigma_only:
        c0 = len(b.chunks)
        b.tiles(lambda kb, nh: ("xyz_encoding_final.weight", nh * 128, 128, 64 * kb, 64, W, False, 128), list(range(nkb)),
                NH, 1, 0, 0)
        boff = b.const("xyz_encoding_final.bias", 0, 1, 0, W, W)
        b.round(L.EPI_LINEAR, W, 0, c0, const_off=boff, save_off=b.save_slot("feat", nkb) if training else L.NONE)
        half = W // 2
        c0 = len(b.chunks)
        for kb in range(nkb):
            img = b.image("extra_encoding.0.weight", 0, half, 64 * kb, 64, W + extra_dim, False, half)
            b.chunk(img, 1, kb, 4, half, 0, init=(kb == 0))
        boff = b.const("extra_encoding.0.bias", 0, 1, 0, half, half)
        aux = b.const("rgb.0.weight", 0, 3, 0, half, half)
        b.const("rgb.0.bias", 0, 1, 0, 3, 3)  # lands at aux + 3*half
        b.round(L.EPI_NERF_RGB, half, 0, c0, raybias=(0 if extra_dim > 0 else -1), const_off=boff, aux_off=aux,
                save_off=b.save_slot("he", _ceil(half, 64)) if training else L.NONE,
                mask_off=b.mask_slot("he", half) if training else L.NONE)
    return b.finish(n_raybias=1 if (extra_dim > 0 and not sigma_only) else 0)


def nof_resident_ok(D: int, W: int, cx: int, skips: Sequence[int]) -> bool:
    """Whether a NoF of these shapes runs on the resident-weight kernel: W = 128, at most one skip layer (its x0 part
    is precomputed into the second half of the slot's accumulator in round 0), forward and backward streams <= 144 KB."""
    if not NOF_RESIDENT or W != 128:
        return False
    n_skip = len([i for i in range(1, D) if i in skips])
    if n_skip > 1:
        return False
    fwd = BLK * (1 + n_skip) + (D - 1) * 2 * BLK + 2 * 2048
    bwd = BLK + (D - 1) * 2 * BLK + (1 + n_skip) * 2 * (BLK // 2)
    return max(fwd, bwd) <= RES_BYTES


def nof_forward_plan(D: int, W: int, cx: int, skips: Sequence[int], extra_dim: int, use_quat: bool,
                     training: bool) -> Plan:
    """models/nof.py:55-85 as a chain program.  The per-ray index-embedding columns of layer 1 and of
    the skip layers are folded into per-ray bias vectors (mcf_ray_bias), in list order.

    Resident-weight form (nof_resident_ok): the first-layer operand x0 shares the activation buffer, so it is gone
    after round 0; the skip layer's x0 part (cat([inputs, h]) @ W^T = inputs @ Wx^T + h @ Wh^T, models/nof.py:71-72)
    is therefore issued in round 0 into accumulator columns [128, 256) and the skip round accumulates its h part on
    top -- the same summation order as the streamed form, so the results are bit-identical."""
    _check_common(W, cx)
    b = _Builder(W)
    nkb, NH = W // 64, W // 128
    resident = nof_resident_ok(D, W, cx, skips)
    if resident and NOF_KERNEL == "ts":
        resident = 2          # x0 stays in tensor memory: the plain program (no round-0 precompute)
    precompute = resident == 1
    kx = _ceil(cx, 16)
    if training:
        b.save_slot("x0", 1)
    rb = 0
    for i in range(D):
        wname, bname = f"nof_encoding_{i+1}.0.weight", f"nof_encoding_{i+1}.0.bias"
        cin = cx + extra_dim
        ld = cin if i == 0 else (W + cin if i in skips else W)
        c0 = len(b.chunks)
        is_skip = i in skips and i > 0
        acc = 128 if (precompute and is_skip) else 0
        for si, (abuf, kb, ks, col0, ncols) in enumerate(_trunk_sources(i, skips, cx, nkb, extra_dim)):
            if precompute and is_skip and abuf == 0:
                continue        # issued in round 0 (below)
            for nh in range(NH):
                img = b.image(wname, nh * 128, 128, col0, ncols, ld, False, 128)
                b.chunk(img, abuf, kb, ks, 128, acc + nh * 128, init=(si == 0))
        if precompute and i == 0:
            for j in range(1, D):
                if j in skips:
                    img = b.image(f"nof_encoding_{j+1}.0.weight", 0, 128, 0, cx, W + cin, False, 128)
                    b.chunk(img, 0, 0, kx, 128, 128, init=True)
        folded = (i == 0 or i in skips) and extra_dim > 0
        boff = 0 if folded else b.const(bname, 0, 1, 0, W, W)
        save = b.save_slot(f"h{i+1}", nkb) if training else L.NONE
        mask = b.mask_slot(f"h{i+1}", W) if training else L.NONE
        b.round(L.EPI_RELU, W, acc, c0, raybias=(rb if folded else -1), const_off=boff, save_off=save, mask_off=mask)
        if folded:
            rb += 1
    n_head = 9 if use_quat else 3
    c0 = len(b.chunks)
    for kb in range(nkb):
        img = b.image("nof_encoding_final.weight", 0, n_head, 64 * kb, 64, W, False, 16)
        b.chunk(img, 1, kb, 4, 16, 0, init=(kb == 0))
    boff = b.const("nof_encoding_final.bias", 0, 1, 0, n_head, n_head, pad_to=16)
    b.round(L.EPI_NOF_HEAD, 16, 0, c0, const_off=boff)
    if rb > 4:
        raise ValueError("at most 4 folded layers (first + 3 skips) are supported")
    return b.finish(n_raybias=rb, kind=1, resident=int(resident))


def folded_layers(D: int, skips: Sequence[int]) -> List[int]:
    """0-based trunk layers of a NoF whose extra-feature columns are folded, in raybias order."""
    return [i for i in range(D) if i == 0 or i in skips]


# ================================================================================================
# backward: dX chain programs and weight-gradient job lists
# ================================================================================================
@dataclass
class GradJob:
    """One mcf_dw_gemm call: staging[i][j] = sum_rows P[row][i] * Q[row][j]."""
    p_src: str      # 'fwd' | 'bwd' save record
    p_off: int
    p_cols: int
    q_src: str
    q_off: int
    q_cols: int
    st_off: int     # float offset into the staging buffer
    ld: int
    n_i: int
    n_j: int
    colsum_off: int  # float offset in staging of colsum_p, or -1
    params: tuple    # parameter names this job feeds
    q_split: int = -1  # leading 64-column Q blocks taken from q_src; the rest from the per-ray feature images
    q2_off: int = 0


@dataclass
class GradPlan:
    jobs: List[GradJob]
    unpack: np.ndarray           # UNPACK_DT entries staging -> flat gradient buffer
    unpack_targets: List[tuple]  # per unpack entry: (parameter name, float offset inside that parameter)
    staging_floats: int
    head_colsum: tuple           # (ncols, stride, staging float offset) of the fp32 head-gradient column sums
    param_names: List[str]
    param_offsets: Dict[str, int]
    param_shapes: Dict[str, tuple]
    total_floats: int


class _GradBuilder:
    def __init__(self, shapes: Dict[str, tuple]):
        self.jobs: List[GradJob] = []
        self.unpack: List[tuple] = []
        self.targets: List[tuple] = []
        self.st = 0
        self.names = list(shapes)
        self.shapes = dict(shapes)
        self.offsets, off = {}, 0
        for n, shp in shapes.items():
            self.offsets[n] = off
            off += _ceil(int(np.prod(shp)), 4) * 4
        self.total = off

    def alloc(self, n: int) -> int:
        off = self.st
        self.st += _ceil(n, 4) * 4
        return off

    def job(self, P, Q, n_i, n_j, params, colsum=False, q_split=-1) -> GradJob:
        n_j4 = _ceil(n_j, 4) * 4
        ld = n_j4
        st = self.alloc(n_i * ld)
        cs = self.alloc(n_i) if colsum else -1
        j = GradJob(P[0], P[1], P[2], Q[0], Q[1], Q[2], st, ld, n_i, n_j4, cs, tuple(params), q_split, 0)
        self.jobs.append(j)
        return j

    def scatter(self, src_off, src_ld, name, row0, col0, nrows, ncols, transposed=False):
        shp = self.shapes[name]
        dst_ld = shp[1] if len(shp) == 2 else shp[0]
        inner = row0 * dst_ld + col0 if len(shp) == 2 else col0
        dst = self.offsets[name] + inner
        self.unpack.append((src_off, dst, src_ld, dst_ld, nrows, ncols, int(transposed), 0))
        self.targets.append((name, inner))

    def finish(self, head_colsum) -> GradPlan:
        return GradPlan(self.jobs, np.array(self.unpack, dtype=L.UNPACK_DT), list(self.targets), max(self.st, 4),
                        head_colsum,
                        self.names, self.offsets, self.shapes, self.total)


def _bwd_trunk(b: _Builder, fwd: Plan, D: int, W: int, cx: int, skips, skip_extra: int, prefix: str,
               need_dx: bool) -> None:
    """Rounds that take dY_D (already in H) down to dY_1 (and the PE gradient)."""
    nkb, NH = W // 64, W // 128
    cin_extra = cx + skip_extra
    for i in range(D - 1, -1, -1):
        wname = f"{prefix}_{i+1}.0.weight"
        is_skip = i in skips and i > 0
        ld = cin_extra if i == 0 else (W + cin_extra if is_skip else W)
        if (i == 0 or is_skip) and need_dx:
            c0 = len(b.chunks)
            for kb in range(nkb):
                img = b.image(wname, 0, cx, 64 * kb, 64, ld, True, 64)
                b.chunk(img, 1, kb, 4, 64, 0, init=(kb == 0))
            b.round(L.EPI_B_DPE, 64, 0, c0, aux_off=(1 if i == 0 else 0))
        if i == 0:
            break
        base = cin_extra if is_skip else 0
        c0 = len(b.chunks)
        b.tiles(lambda kb, nh: (wname, base + nh * 128, 128, 64 * kb, 64, ld, True, 128), list(range(nkb)), NH, 1, 0, 0)
        b.round(L.EPI_B_MASK, W, 0, c0, save_off=b.save_slot(f"dy{i}", nkb), mask_off=fwd.offsets[f"mask_h{i}"])


def job_table(gp: "GradPlan", wanted: set) -> np.ndarray:
    """Device job table (mcf_dw_job_t) of a gradient plan; jobs feeding no wanted parameter are disabled."""
    rows = []
    src = {"fwd": 0, "bwd": 1, "aux": 2}   # forward / backward save record, per-ray feature images (mcf_rayfeat_image)
    for j in gp.jobs:
        rows.append((j.p_off, j.q_off, src[j.p_src], src[j.q_src], j.p_cols, j.q_cols,
                     j.st_off, j.ld, j.n_i, j.n_j, j.colsum_off, int(any(n in wanted for n in j.params)),
                     j.q_split, j.q2_off))
    return np.array(rows, dtype=L.DWJOB_DT)


def nerf_backward_plan(D: int, W: int, cx: int, skips: Sequence[int], extra_dim: int, need_dx: bool,
                       fwd: Plan) -> Plan:
    """dX chain of models/nerf.py:61-102 (autograd of the forward program)."""
    _check_common(W, cx)
    if W != 256:
        raise ValueError("NeRF training through the fused kernels requires W == 256")
    b = _Builder(W)
    nkb, NH, half = W // 64, W // 128, W // 2
    b.save_slot("dhead", 1)
    b.save_slot("dye", _ceil(half, 64))
    # round 0: through extra_encoding (feat part)
    c0 = len(b.chunks)
    b.tiles(lambda kb, nh: ("extra_encoding.0.weight", nh * 128, 128, 64 * kb, 64, W + extra_dim, True, 128),
            list(range(_ceil(half, 64))), NH, 1, 0, 0)
    wrgb = b.const("rgb.0.weight", 0, 3, 0, half, half)
    b.round(L.EPI_B_LINEAR, W, 0, c0, aux_off=wrgb, save_off=b.save_slot("dyf", nkb),
            mask_off=fwd.offsets["mask_he"])
    # round 1: through xyz_encoding_final, add the sigma head, mask with h_D
    c0 = len(b.chunks)
    b.tiles(lambda kb, nh: ("xyz_encoding_final.weight", nh * 128, 128, 64 * kb, 64, W, True, 128), list(range(nkb)), NH,
            1, 0, 0)
    wsig = b.const("sigma.weight", 0, 1, 0, W, W)
    b.round(L.EPI_B_MASK_SIGMA, W, 0, c0, aux_off=wsig, save_off=b.save_slot(f"dy{D}", nkb),
            mask_off=fwd.offsets[f"mask_h{D}"])
    _bwd_trunk(b, fwd, D, W, cx, tuple(skips), 0, "xyz_encoding", need_dx)
    return b.finish()


def nof_backward_plan(D: int, W: int, cx: int, skips: Sequence[int], extra_dim: int, use_quat: bool,
                      need_dx: bool, fwd: Plan) -> Plan:
    """dX chain of models/nof.py:55-85."""
    _check_common(W, cx)
    b = _Builder(W)
    nkb, NH = W // 64, W // 128
    n_head = 9 if use_quat else 3
    b.save_slot("ghead", 1)
    c0 = len(b.chunks)
    for nh in range(NH):
        img = b.image("nof_encoding_final.weight", nh * 128, 128, 0, n_head, W, True, 128)
        b.chunk(img, 1, 0, 1, 128, nh * 128, init=True)
    b.round(L.EPI_B_MASK, W, 0, c0, save_off=b.save_slot(f"dy{D}", nkb), mask_off=fwd.offsets[f"mask_h{D}"])
    _bwd_trunk(b, fwd, D, W, cx, tuple(skips), extra_dim, "nof_encoding", need_dx)
    return b.finish(kind=1, resident=fwd.resident)


def nerf_grad_plan(D: int, W: int, cx: int, skips: Sequence[int], extra_dim: int, shapes: Dict[str, tuple],
                   fwd: Plan, bwd: Plan) -> GradPlan:
    g = _GradBuilder(shapes)
    half = W // 2
    F = lambda key, cols: ("fwd", fwd.offsets["save_" + key], cols)
    B = lambda key, cols: ("bwd", bwd.offsets["save_" + key], cols)
    for i in range(D):
        wn, bn = f"xyz_encoding_{i+1}.0.weight", f"xyz_encoding_{i+1}.0.bias"
        P = B(f"dy{i+1}", W)
        is_skip = i in skips and i > 0
        first = True
        if i == 0 or is_skip:
            j = g.job(P, F("x0", 64), W, 64, (wn, bn), colsum=True)
            g.scatter(j.st_off, j.ld, wn, 0, 0, W, cx)
            g.scatter(j.colsum_off, W, bn, 0, 0, 1, W)
            first = False
        if i > 0:
            j = g.job(P, F(f"h{i}", W), W, W, (wn, bn), colsum=first)
            g.scatter(j.st_off, j.ld, wn, 0, cx if is_skip else 0, W, W)
            if first:
                g.scatter(j.colsum_off, W, bn, 0, 0, 1, W)
    j = g.job(B("dyf", W), F(f"h{D}", W), W, W, ("xyz_encoding_final.weight", "xyz_encoding_final.bias"), colsum=True)
    g.scatter(j.st_off, j.ld, "xyz_encoding_final.weight", 0, 0, W, W)
    g.scatter(j.colsum_off, W, "xyz_encoding_final.bias", 0, 0, 1, W)
    en, eb = "extra_encoding.0.weight", "extra_encoding.0.bias"
    j = g.job(B("dye", half), F("feat", W), half, W, (en, eb), colsum=True)
    g.scatter(j.st_off, j.ld, en, 0, 0, half, W)
    g.scatter(j.colsum_off, half, eb, 0, 0, 1, half)
    if extra_dim > 0:
        j = g.job(B("dye", half), ("aux", 0, 64), half, 64, (en,))
        g.scatter(j.st_off, j.ld, en, 0, W, half, extra_dim)
    j = g.job(F("he", half), B("dhead", 64), half, 4, ("rgb.0.weight",))
    g.scatter(j.st_off, j.ld, "rgb.0.weight", 0, 0, 3, half, transposed=True)
    j = g.job(F(f"h{D}", W), B("dhead", 64), W, 4, ("sigma.weight",))
    g.scatter(j.st_off + 3, j.ld, "sigma.weight", 0, 0, 1, W, transposed=True)
    hc = g.alloc(4)
    g.scatter(hc, 4, "rgb.0.bias", 0, 0, 1, 3)
    g.scatter(hc + 3, 4, "sigma.bias", 0, 0, 1, 1)
    return g.finish((4, 4, hc))


def nof_grad_plan(D: int, W: int, cx: int, skips: Sequence[int], extra_dim: int, use_quat: bool,
                  shapes: Dict[str, tuple], fwd: Plan, bwd: Plan) -> GradPlan:
    g = _GradBuilder(shapes)
    n_head = 9 if use_quat else 3
    F = lambda key, cols: ("fwd", fwd.offsets["save_" + key], cols)
    B = lambda key, cols: ("bwd", bwd.offsets["save_" + key], cols)
    for i in range(D):
        wn, bn = f"nof_encoding_{i+1}.0.weight", f"nof_encoding_{i+1}.0.bias"
        P = B(f"dy{i+1}", W)
        is_skip = i in skips and i > 0
        first = True
        if i == 0 or is_skip:
            # Q = [x0 block (forward save record) | per-ray feature block (shared images, mcf_rayfeat_image)]
            j = g.job(P, F("x0", 128), W, 128, (wn, bn), colsum=True, q_split=1)
            g.scatter(j.st_off, j.ld, wn, 0, 0, W, cx)
            if extra_dim > 0:
                g.scatter(j.st_off + 64, j.ld, wn, 0, cx, W, extra_dim)
            g.scatter(j.colsum_off, W, bn, 0, 0, 1, W)
            first = False
        if i > 0:
            j = g.job(P, F(f"h{i}", W), W, W, (wn, bn), colsum=first)
            g.scatter(j.st_off, j.ld, wn, 0, (cx + extra_dim) if is_skip else 0, W, W)
            if first:
                g.scatter(j.colsum_off, W, bn, 0, 0, 1, W)
    j = g.job(F(f"h{D}", W), B("ghead", 64), W, 12, ("nof_encoding_final.weight",))
    g.scatter(j.st_off, j.ld, "nof_encoding_final.weight", 0, 0, n_head, W, transposed=True)
    hc = g.alloc(12)
    g.scatter(hc, 12, "nof_encoding_final.bias", 0, 0, 1, n_head)
    return g.finish((n_head, 12, hc))
