// Host-side builders of the layer programs (the tables mcf_pack / mcf_chain_launch / mcf_dw_gemm_batch /
// mcf_unpack consume) for the two network families of the path, so that a host other than the Python shim can
// drive the MLP entries of the C ABI.  Pure host code; no CUDA calls.
//
// Network shapes: models/nerf.py:28-59 (trunk of D Linear+ReLU layers of width W with skip concatenation
// cat([input_xyz, h]) :85-86, sigma head, xyz_encoding_final, extra_encoding (W + extra -> W/2), rgb head) and
// models/nof.py:40-53 (trunk with cat([inputs, h]) :71-72 where inputs = [xyz encoding | per-ray index encoding],
// 9- or 3-wide head).  The per-ray columns are folded into per-ray biases (mcf_ray_bias), in order of the layers.
// moco_flow_b200/plans.py builds the same tables; tests/test_host_cpu.py holds the two to each other entry by entry.
#include <stdint.h>
#include <string.h>

#include "../../include/moco_flow_b200.h"

namespace {

constexpr uint32_t kBlk = MCF_BLOCK_BYTES;
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kResBytes = 147456;

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

struct Builder {
  mcf_plan_t* p;
  int err = 0;
  bool pair_layout = false;   // width 256: tiles of a layer laid out half by half (see mcf_plan_spec_t.no_pair_merge)

  explicit Builder(mcf_plan_t* out, int width, bool pair = false) : p(out), pair_layout(pair && width == 256) {
    memset(p, 0, sizeof(*p));
    p->width = width;
    const uint32_t none = kNone;
    p->save_x0 = p->save_feat = p->save_he = p->mask_he = none;
    p->save_dhead = p->save_dye = p->save_dyf = p->save_ghead = none;
    for (int i = 0; i < MCF_PLAN_MAX_LAYERS; ++i) p->save_h[i] = p->mask_h[i] = p->save_dy[i] = none;
  }
  int tensor(int id) {
    for (int i = 0; i < p->n_tensors; ++i)
      if (p->tensor_ids[i] == id) return i;
    if (p->n_tensors >= MCF_MAX_PACK_TENSORS) { err = MCF_ERR_UNSUPPORTED; return 0; }
    p->tensor_ids[p->n_tensors] = id;
    return p->n_tensors++;
  }
  void push_pack(uint32_t dst, uint32_t bytes, int kind, int id, int row0, int nrows, int col0, int ncols, int ld, int tr) {
    if (p->n_pack >= MCF_PLAN_MAX_PACK) { err = MCF_ERR_UNSUPPORTED; return; }
    mcf_pack_t& e = p->pack[p->n_pack++];
    e.dst_off = dst; e.bytes = bytes; e.kind = kind; e.tensor = tensor(id); e.row0 = row0; e.nrows = nrows;
    e.col0 = col0; e.ncols = ncols; e.ld = ld; e.transposed = tr;
  }
  // src[row0:row0+nrows, col0:col0+ncols] (row-major, tightly) -> consts; returns the float offset
  uint32_t konst(int id, int row0, int nrows, int col0, int ncols, int ld, int pad_to = 0) {
    const int n = nrows * ncols;
    const uint32_t total = (uint32_t)ceil_div(n > pad_to ? n : pad_to, 4) * 4u;
    const uint32_t off = p->n_consts;
    push_pack(off, total, 1, id, row0, nrows, col0, ncols, ld, 0);
    p->n_consts += total;
    return off;
  }
  struct Img { uint32_t off, bytes; };
  Img image(int id, int row0, int nrows, int col0, int ncols, int ld, bool transposed, int rows_padded) {
    Img im = {p->wpack_bytes, (uint32_t)rows_padded * 128u};
    push_pack(im.off, im.bytes, 0, id, row0, nrows, col0, ncols, ld, transposed ? 1 : 0);
    p->wpack_bytes += im.bytes;
    return im;
  }
  void chunk(Img im, int a_buf, int a_kblock, int ksteps, int n, int acc_col, bool init) {
    if (p->n_chunks >= MCF_PLAN_MAX_CHUNKS) { err = MCF_ERR_UNSUPPORTED; return; }
    mcf_chunk_t& c = p->chunks[p->n_chunks++];
    c.src_off = im.off; c.bytes = im.bytes; c.a_buf = (uint8_t)a_buf; c.a_kblock = (uint8_t)a_kblock;
    c.ksteps = (uint8_t)ksteps; c.flags = init ? 1 : 0; c.n = (uint16_t)n; c.acc_col = (uint16_t)acc_col;
    // two consecutive 128-row halves of one [256 x 64] weight tile in an (even, odd) pair of ring stages: one
    // contiguous 32 KB K-major tile -> the kernel issues N = 256 instructions (bit 1 of the first chunk's flags)
    if (p->n_chunks >= 2 && p->n_chunks % 2 == 0) {
      mcf_chunk_t& a = p->chunks[p->n_chunks - 2];
      mcf_chunk_t& b = p->chunks[p->n_chunks - 1];
      if (a.a_buf == b.a_buf && a.a_kblock == b.a_kblock && a.ksteps == b.ksteps && (a.flags & 1) == (b.flags & 1) &&
          a.n == b.n && a.n == 128 && a.bytes == kBlk && b.bytes == kBlk && b.acc_col == a.acc_col + 128 &&
          (b.src_off == a.src_off + kBlk || pair_layout))
        a.flags |= 2;
    }
  }
  // the [NH*128 x 64] weight tiles of k-blocks 0..nkb-1 of one layer: images in stream order, chunks in consumption
  // order (k-block major, then the 128-row halves)
  template <typename F>
  void tiles(F image_of, int nkb, int NH, int a_buf, int acc0, int init_kb) {
    Img imgs[8][2];
    if (pair_layout && NH == 2) {
      for (int nh = 0; nh < NH; ++nh)
        for (int kb = 0; kb < nkb; ++kb) imgs[kb][nh] = image_of(kb, nh);
    } else {
      for (int kb = 0; kb < nkb; ++kb)
        for (int nh = 0; nh < NH; ++nh) imgs[kb][nh] = image_of(kb, nh);
    }
    for (int kb = 0; kb < nkb; ++kb)
      for (int nh = 0; nh < NH; ++nh) chunk(imgs[kb][nh], a_buf, kb, 4, 128, acc0 + nh * 128, kb == init_kb);
  }
  void round(int epi, int n_out, int acc_col, int chunk_begin, int raybias = -1, uint32_t const_off = 0,
             uint32_t aux_off = 0, uint32_t save_off = kNone, uint32_t mask_off = kNone) {
    if (p->n_rounds >= MCF_PLAN_MAX_ROUNDS) { err = MCF_ERR_UNSUPPORTED; return; }
    // flag bit 3: one CTA's halves of this pair tile and of the next k-block's are adjacent in the stream
    for (int c = chunk_begin; c + 3 < p->n_chunks; ++c) {
      mcf_chunk_t &a = p->chunks[c], &a1 = p->chunks[c + 1], &b = p->chunks[c + 2], &b1 = p->chunks[c + 3];
      if ((a.flags & 2) && (b.flags & 2) && (c - chunk_begin) % 2 == 0 && a.a_buf == b.a_buf && b.a_kblock == a.a_kblock + 1 &&
          a.ksteps == 4 && b.ksteps == 4 && !(b.flags & 1) && b.src_off == a.src_off + kBlk && b1.src_off == a1.src_off + kBlk)
        a.flags |= 8;
    }
    mcf_round_t& r = p->rounds[p->n_rounds++];
    r.epi = (uint16_t)epi; r.n_out = (uint16_t)n_out; r.acc_col = (uint16_t)acc_col;
    r.chunk_begin = (uint16_t)chunk_begin; r.chunk_end = (uint16_t)p->n_chunks; r.raybias = (int16_t)raybias;
    r.const_off = const_off; r.aux_off = aux_off; r.save_off = save_off; r.mask_off = mask_off; r.reserved = 0;
  }
  uint32_t save_slot(uint32_t* named, int n_blocks) {
    const uint32_t off = p->save_tile_bytes;
    *named = off;
    p->save_tile_bytes += (uint32_t)n_blocks * kBlk;
    return off;
  }
  uint32_t mask_slot(uint32_t* named, int n_cols) {
    const uint32_t off = p->mask_tile_words;
    *named = off;
    p->mask_tile_words += (uint32_t)ceil_div(n_cols, 32) * 128u;
    return off;
  }
  int finish(int n_raybias, int kind, int resident) {
    if (p->n_consts < 4) p->n_consts = 4;
    p->n_raybias = n_raybias; p->kind = kind; p->resident = resident;
    if (resident && p->wpack_bytes > kResBytes) err = MCF_ERR_UNSUPPORTED;
    return err;
  }
};

bool is_skip(const mcf_plan_spec_t& s, int i) {
  for (int k = 0; k < s.n_skips; ++k)
    if (s.skips[k] == i) return true;
  return false;
}

// canonical parameter ids (module definition order = nn.Module.named_parameters order)
inline int id_trunk_w(int i) { return 2 * i; }
inline int id_trunk_b(int i) { return 2 * i + 1; }
inline int id_final_w(const mcf_plan_spec_t& s) { return 2 * s.D; }
inline int id_final_b(const mcf_plan_spec_t& s) { return 2 * s.D + 1; }
inline int id_extra_w(const mcf_plan_spec_t& s) { return 2 * s.D + 2; }
inline int id_extra_b(const mcf_plan_spec_t& s) { return 2 * s.D + 3; }
inline int id_sigma_w(const mcf_plan_spec_t& s) { return 2 * s.D + 4; }
inline int id_sigma_b(const mcf_plan_spec_t& s) { return 2 * s.D + 5; }
inline int id_rgb_w(const mcf_plan_spec_t& s) { return 2 * s.D + 6; }
inline int id_rgb_b(const mcf_plan_spec_t& s) { return 2 * s.D + 7; }

int check_spec(const mcf_plan_spec_t& s) {
  if (s.W != 128 && s.W != 256) return MCF_ERR_UNSUPPORTED;
  if (s.cx < 1 || s.cx > 64 || s.D < 1 || s.D > MCF_PLAN_MAX_LAYERS - 1 || s.n_skips < 0 || s.n_skips > 8 || s.extra_dim < 0 ||
      s.extra_dim > 64)
    return MCF_ERR_BAD_ARG;
  if (s.family != 0 && s.family != 1) return MCF_ERR_BAD_ARG;
  return 0;
}

bool nof_resident_ok(const mcf_plan_spec_t& s) {
  if (s.nof_kernel == 0 || s.W != 128) return false;
  int n_skip = 0;
  for (int i = 1; i < s.D; ++i) n_skip += is_skip(s, i) ? 1 : 0;
  if (n_skip > 1) return false;
  const uint32_t fwd = kBlk * (1 + n_skip) + (uint32_t)(s.D - 1) * 2 * kBlk + 2 * 2048;
  const uint32_t bwd = kBlk + (uint32_t)(s.D - 1) * 2 * kBlk + (uint32_t)(1 + n_skip) * 2 * (kBlk / 2);
  return (fwd > bwd ? fwd : bwd) <= kResBytes;
}

int nerf_forward(const mcf_plan_spec_t& s, mcf_plan_t* out) {
  Builder b(out, s.W, !s.no_pair_merge);
  const int W = s.W, D = s.D, cx = s.cx, nkb = W / 64, NH = W / 128, kx = ceil_div(cx, 16);
  if (s.training) b.save_slot(&out->save_x0, 1);
  for (int i = 0; i < D; ++i) {
    const bool skip = is_skip(s, i);
    const int ld = i == 0 ? cx : (skip ? W + cx : W);
    const int c0 = out->n_chunks;
    const bool has_x0 = i == 0 || skip;
    if (has_x0)
      for (int nh = 0; nh < NH; ++nh) b.chunk(b.image(id_trunk_w(i), nh * 128, 128, 0, cx, ld, false, 128), 0, 0, kx, 128, nh * 128, true);
    if (i > 0) {
      const int base = skip ? cx : 0;
      b.tiles([&](int kb, int nh) { return b.image(id_trunk_w(i), nh * 128, 128, base + 64 * kb, 64, ld, false, 128); }, nkb, NH, 1,
              0, has_x0 ? -1 : 0);
    }
    const uint32_t boff = b.konst(id_trunk_b(i), 0, 1, 0, W, W);
    const bool last = i == D - 1;
    uint32_t aux = 0;
    if (last) {
      aux = b.konst(id_sigma_w(s), 0, 1, 0, W, W);
      b.konst(id_sigma_b(s), 0, 1, 0, 1, 1);   // lands at aux + W
    }
    const uint32_t save = s.training ? b.save_slot(&out->save_h[i + 1], nkb) : kNone;
    const uint32_t mask = s.training ? b.mask_slot(&out->mask_h[i + 1], W) : kNone;
    b.round(last ? MCF_EPI_RELU_SIGMA : MCF_EPI_RELU, W, 0, c0, -1, boff, aux, save, mask);
  }
  if (!s.sigma_only) {
    int c0 = out->n_chunks;
    b.tiles([&](int kb, int nh) { return b.image(id_final_w(s), nh * 128, 128, 64 * kb, 64, W, false, 128); }, nkb, NH, 1, 0, 0);
    uint32_t boff = b.konst(id_final_b(s), 0, 1, 0, W, W);
    b.round(MCF_EPI_LINEAR, W, 0, c0, -1, boff, 0, s.training ? b.save_slot(&out->save_feat, nkb) : kNone);
    const int half = W / 2;
    c0 = out->n_chunks;
    for (int kb = 0; kb < nkb; ++kb) b.chunk(b.image(id_extra_w(s), 0, half, 64 * kb, 64, W + s.extra_dim, false, half), 1, kb, 4, half, 0, kb == 0);
    boff = b.konst(id_extra_b(s), 0, 1, 0, half, half);
    const uint32_t aux = b.konst(id_rgb_w(s), 0, 3, 0, half, half);
    b.konst(id_rgb_b(s), 0, 1, 0, 3, 3);       // lands at aux + 3*half
    const uint32_t save = s.training ? b.save_slot(&out->save_he, ceil_div(half, 64)) : kNone;
    const uint32_t mask = s.training ? b.mask_slot(&out->mask_he, half) : kNone;
    b.round(MCF_EPI_NERF_RGB, half, 0, c0, s.extra_dim > 0 ? 0 : -1, boff, aux, save, mask);
  }
  return b.finish((s.extra_dim > 0 && !s.sigma_only) ? 1 : 0, 0, 0);
}

int nof_forward(const mcf_plan_spec_t& s, mcf_plan_t* out) {
  Builder b(out, s.W);
  const int W = s.W, D = s.D, cx = s.cx, nkb = W / 64, NH = W / 128, kx = ceil_div(cx, 16);
  int resident = nof_resident_ok(s) ? (s.nof_kernel == 2 ? 2 : 1) : 0;
  const bool precompute = resident == 1;
  if (s.training) b.save_slot(&out->save_x0, 1);
  int rb = 0;
  const int cin = cx + s.extra_dim;
  for (int i = 0; i < D; ++i) {
    const bool skip = is_skip(s, i) && i > 0;
    const int ld = i == 0 ? cin : (is_skip(s, i) ? W + cin : W);
    const int c0 = out->n_chunks;
    const int acc = (precompute && skip) ? 128 : 0;
    int si = 0;
    if (i == 0 || is_skip(s, i)) {
      if (!(precompute && skip))
        for (int nh = 0; nh < NH; ++nh) b.chunk(b.image(id_trunk_w(i), nh * 128, 128, 0, cx, ld, false, 128), 0, 0, kx, 128, acc + nh * 128, si == 0);
      ++si;
    }
    if (i > 0) {
      const int base = is_skip(s, i) ? cin : 0;
      for (int kb = 0; kb < nkb; ++kb, ++si)
        for (int nh = 0; nh < NH; ++nh)
          b.chunk(b.image(id_trunk_w(i), nh * 128, 128, base + 64 * kb, 64, ld, false, 128), 1, kb, 4, 128, acc + nh * 128, si == 0);
    }
    if (precompute && i == 0)
      for (int j = 1; j < D; ++j)
        if (is_skip(s, j)) b.chunk(b.image(id_trunk_w(j), 0, 128, 0, cx, W + cin, false, 128), 0, 0, kx, 128, 128, true);
    const bool folded = (i == 0 || is_skip(s, i)) && s.extra_dim > 0;
    const uint32_t boff = folded ? 0 : b.konst(id_trunk_b(i), 0, 1, 0, W, W);
    const uint32_t save = s.training ? b.save_slot(&out->save_h[i + 1], nkb) : kNone;
    const uint32_t mask = s.training ? b.mask_slot(&out->mask_h[i + 1], W) : kNone;
    b.round(MCF_EPI_RELU, W, acc, c0, folded ? rb : -1, boff, 0, save, mask);
    if (folded) ++rb;
  }
  const int n_head = s.use_quat ? 9 : 3;
  const int c0 = out->n_chunks;
  for (int kb = 0; kb < nkb; ++kb) b.chunk(b.image(id_final_w(s), 0, n_head, 64 * kb, 64, W, false, 16), 1, kb, 4, 16, 0, kb == 0);
  const uint32_t boff = b.konst(id_final_b(s), 0, 1, 0, n_head, n_head, 16);
  b.round(MCF_EPI_NOF_HEAD, 16, 0, c0, -1, boff);
  if (rb > 4) return MCF_ERR_UNSUPPORTED;
  return b.finish(rb, 1, resident);
}

// rounds that take dY_D (already the A operand) down to dY_1 (and the encoder gradient)
void bwd_trunk(Builder& b, const mcf_plan_spec_t& s, const mcf_plan_t& fwd, int skip_extra) {
  mcf_plan_t* out = b.p;
  const int W = s.W, D = s.D, cx = s.cx, nkb = W / 64, NH = W / 128;
  const int cin_extra = cx + skip_extra;
  for (int i = D - 1; i >= 0; --i) {
    const bool skip = is_skip(s, i) && i > 0;
    const int ld = i == 0 ? cin_extra : (skip ? W + cin_extra : W);
    if ((i == 0 || skip) && s.need_dx) {
      const int c0 = out->n_chunks;
      for (int kb = 0; kb < nkb; ++kb) b.chunk(b.image(id_trunk_w(i), 0, cx, 64 * kb, 64, ld, true, 64), 1, kb, 4, 64, 0, kb == 0);
      b.round(MCF_EPI_B_DPE, 64, 0, c0, -1, 0, i == 0 ? 1u : 0u);
    }
    if (i == 0) break;
    const int base = skip ? cin_extra : 0;
    const int c0 = out->n_chunks;
    b.tiles([&](int kb, int nh) { return b.image(id_trunk_w(i), base + nh * 128, 128, 64 * kb, 64, ld, true, 128); }, nkb, NH, 1, 0, 0);
    b.round(MCF_EPI_B_MASK, W, 0, c0, -1, 0, 0, b.save_slot(&out->save_dy[i], nkb), fwd.mask_h[i]);
  }
}

int nerf_backward(const mcf_plan_spec_t& s, const mcf_plan_t& fwd, mcf_plan_t* out) {
  if (s.W != 256) return MCF_ERR_UNSUPPORTED;
  Builder b(out, s.W, !s.no_pair_merge);
  const int W = s.W, D = s.D, nkb = W / 64, NH = W / 128, half = W / 2;
  b.save_slot(&out->save_dhead, 1);
  b.save_slot(&out->save_dye, ceil_div(half, 64));
  int c0 = out->n_chunks;
  b.tiles([&](int kb, int nh) { return b.image(id_extra_w(s), nh * 128, 128, 64 * kb, 64, W + s.extra_dim, true, 128); },
          ceil_div(half, 64), NH, 1, 0, 0);
  const uint32_t wrgb = b.konst(id_rgb_w(s), 0, 3, 0, half, half);
  b.round(MCF_EPI_B_LINEAR, W, 0, c0, -1, 0, wrgb, b.save_slot(&out->save_dyf, nkb), fwd.mask_he);
  c0 = out->n_chunks;
  b.tiles([&](int kb, int nh) { return b.image(id_final_w(s), nh * 128, 128, 64 * kb, 64, W, true, 128); }, nkb, NH, 1, 0, 0);
  const uint32_t wsig = b.konst(id_sigma_w(s), 0, 1, 0, W, W);
  b.round(MCF_EPI_B_MASK_SIGMA, W, 0, c0, -1, 0, wsig, b.save_slot(&out->save_dy[D], nkb), fwd.mask_h[D]);
  bwd_trunk(b, s, fwd, 0);
  return b.finish(0, 0, 0);
}

int nof_backward(const mcf_plan_spec_t& s, const mcf_plan_t& fwd, mcf_plan_t* out) {
  Builder b(out, s.W);
  const int W = s.W, D = s.D, nkb = W / 64, NH = W / 128;
  const int n_head = s.use_quat ? 9 : 3;
  b.save_slot(&out->save_ghead, 1);
  const int c0 = out->n_chunks;
  for (int nh = 0; nh < NH; ++nh) b.chunk(b.image(id_final_w(s), nh * 128, 128, 0, n_head, W, true, 128), 1, 0, 1, 128, nh * 128, true);
  b.round(MCF_EPI_B_MASK, W, 0, c0, -1, 0, 0, b.save_slot(&out->save_dy[D], nkb), fwd.mask_h[D]);
  bwd_trunk(b, s, fwd, s.extra_dim);
  return b.finish(0, 1, fwd.resident);
}

// ------------------------------------------------------------------------------------------------
// weight-gradient jobs and the scatter of their results into the flat gradient layout
// ------------------------------------------------------------------------------------------------
struct GradBuilder {
  mcf_grad_plan_t* g;
  const mcf_plan_spec_t& s;
  int err = 0;
  int rows[MCF_MAX_PACK_TENSORS], cols[MCF_MAX_PACK_TENSORS];   // parameter shapes (bias: rows = 1)

  GradBuilder(mcf_grad_plan_t* out, const mcf_plan_spec_t& spec) : g(out), s(spec) {
    memset(g, 0, sizeof(*g));
    const int W = s.W, D = s.D;
    int n = 0;
    auto param = [&](int r, int c) { rows[n] = r; cols[n] = c; ++n; };
    const int cin = s.cx + (s.family == 1 ? s.extra_dim : 0);
    for (int i = 0; i < D; ++i) {
      param(W, i == 0 ? cin : (is_skip(s, i) ? W + cin : W));
      param(1, W);
    }
    if (s.family == 0) {
      param(W, W); param(1, W);
      param(W / 2, W + s.extra_dim); param(1, W / 2);
      param(1, W); param(1, 1);
      param(3, W / 2); param(1, 3);
    } else {
      const int nh = s.use_quat ? 9 : 3;
      param(nh, W); param(1, nh);
    }
    g->n_params = n;
    uint32_t off = 0;
    for (int i = 0; i < n; ++i) {
      g->param_offset[i] = off;
      off += (uint32_t)ceil_div(rows[i] * cols[i], 4) * 4u;
    }
    g->total_floats = off;
  }
  uint32_t alloc(int n) {
    const uint32_t off = g->staging_floats;
    g->staging_floats += (uint32_t)ceil_div(n, 4) * 4u;
    return off;
  }
  struct Src { int src; uint32_t off; int cols; };
  mcf_dw_job_t* job(Src P, Src Q, int n_i, int n_j, int p0, int p1, bool colsum, int q_split = -1) {
    if (g->n_jobs >= MCF_PLAN_MAX_JOBS) { err = MCF_ERR_UNSUPPORTED; return &g->jobs[0]; }
    const int n_j4 = ceil_div(n_j, 4) * 4;
    mcf_dw_job_t& j = g->jobs[g->n_jobs];
    g->job_params[g->n_jobs][0] = p0;
    g->job_params[g->n_jobs][1] = p1;
    ++g->n_jobs;
    j.p_off = P.off; j.q_off = Q.off; j.p_src = P.src; j.q_src = Q.src; j.p_cols = P.cols; j.q_cols = Q.cols;
    j.st_off = alloc(n_i * n_j4); j.ld = n_j4; j.n_i = n_i; j.n_j = n_j4;
    j.colsum_off = colsum ? (int32_t)alloc(n_i) : -1;
    j.enabled = 1; j.q_split = q_split; j.q2_off = 0;
    return &j;
  }
  void scatter(uint32_t src_off, int src_ld, int param, int row0, int col0, int nrows, int ncols, bool transposed = false) {
    if (g->n_unpack >= MCF_MAX_UNPACK_PTRS) { err = MCF_ERR_UNSUPPORTED; return; }
    const bool matrix = rows[param] > 1 || param % 2 == 0;   // weights are 2-D (also the 1 x W sigma weight)
    const int dst_ld = matrix ? cols[param] : cols[param];
    const uint32_t inner = matrix ? (uint32_t)(row0 * dst_ld + col0) : (uint32_t)col0;
    mcf_unpack_t& u = g->unpack[g->n_unpack];
    u.src_off = src_off; u.dst_off = g->param_offset[param] + inner; u.src_ld = src_ld; u.dst_ld = dst_ld;
    u.nrows = nrows; u.ncols = ncols; u.transposed = transposed ? 1 : 0; u.reserved = 0;
    g->unpack_param[g->n_unpack] = param;
    g->unpack_inner[g->n_unpack] = inner;
    ++g->n_unpack;
  }
};

int nerf_gradients(const mcf_plan_spec_t& s, const mcf_plan_t& fwd, const mcf_plan_t& bwd, mcf_grad_plan_t* out) {
  GradBuilder g(out, s);
  const int W = s.W, D = s.D, cx = s.cx, half = W / 2;
  typedef GradBuilder::Src Src;
  for (int i = 0; i < D; ++i) {
    const int wn = id_trunk_w(i), bn = id_trunk_b(i);
    const Src P = {1, bwd.save_dy[i + 1], W};
    const bool skip = is_skip(s, i) && i > 0;
    bool first = true;
    if (i == 0 || skip) {
      mcf_dw_job_t* j = g.job(P, Src{0, fwd.save_x0, 64}, W, 64, wn, bn, true);
      g.scatter(j->st_off, j->ld, wn, 0, 0, W, cx);
      g.scatter((uint32_t)j->colsum_off, W, bn, 0, 0, 1, W);
      first = false;
    }
    if (i > 0) {
      mcf_dw_job_t* j = g.job(P, Src{0, fwd.save_h[i], W}, W, W, wn, bn, first);
      g.scatter(j->st_off, j->ld, wn, 0, skip ? cx : 0, W, W);
      if (first) g.scatter((uint32_t)j->colsum_off, W, bn, 0, 0, 1, W);
    }
  }
  mcf_dw_job_t* j = g.job(Src{1, bwd.save_dyf, W}, Src{0, fwd.save_h[D], W}, W, W, id_final_w(s), id_final_b(s), true);
  g.scatter(j->st_off, j->ld, id_final_w(s), 0, 0, W, W);
  g.scatter((uint32_t)j->colsum_off, W, id_final_b(s), 0, 0, 1, W);
  j = g.job(Src{1, bwd.save_dye, half}, Src{0, fwd.save_feat, W}, half, W, id_extra_w(s), id_extra_b(s), true);
  g.scatter(j->st_off, j->ld, id_extra_w(s), 0, 0, half, W);
  g.scatter((uint32_t)j->colsum_off, half, id_extra_b(s), 0, 0, 1, half);
  if (s.extra_dim > 0) {
    j = g.job(Src{1, bwd.save_dye, half}, Src{2, 0, 64}, half, 64, id_extra_w(s), -1, false);
    g.scatter(j->st_off, j->ld, id_extra_w(s), 0, W, half, s.extra_dim);
  }
  j = g.job(Src{0, fwd.save_he, half}, Src{1, bwd.save_dhead, 64}, half, 4, id_rgb_w(s), -1, false);
  g.scatter(j->st_off, j->ld, id_rgb_w(s), 0, 0, 3, half, true);
  j = g.job(Src{0, fwd.save_h[D], W}, Src{1, bwd.save_dhead, 64}, W, 4, id_sigma_w(s), -1, false);
  g.scatter(j->st_off + 3, j->ld, id_sigma_w(s), 0, 0, 1, W, true);
  const uint32_t hc = g.alloc(4);
  g.scatter(hc, 4, id_rgb_b(s), 0, 0, 1, 3);
  g.scatter(hc + 3, 4, id_sigma_b(s), 0, 0, 1, 1);
  out->head_ncols = 4; out->head_stride = 4; out->head_off = hc;
  if (out->staging_floats < 4) out->staging_floats = 4;
  return g.err;
}

int nof_gradients(const mcf_plan_spec_t& s, const mcf_plan_t& fwd, const mcf_plan_t& bwd, mcf_grad_plan_t* out) {
  GradBuilder g(out, s);
  const int W = s.W, D = s.D, cx = s.cx;
  const int n_head = s.use_quat ? 9 : 3;
  typedef GradBuilder::Src Src;
  for (int i = 0; i < D; ++i) {
    const int wn = id_trunk_w(i), bn = id_trunk_b(i);
    const Src P = {1, bwd.save_dy[i + 1], W};
    const bool skip = is_skip(s, i) && i > 0;
    bool first = true;
    if (i == 0 || skip) {
      // Q = [x0 block (forward save record) | per-ray feature block (shared images, mcf_rayfeat_image)]
      mcf_dw_job_t* j = g.job(P, Src{0, fwd.save_x0, 128}, W, 128, wn, bn, true, 1);
      g.scatter(j->st_off, j->ld, wn, 0, 0, W, cx);
      if (s.extra_dim > 0) g.scatter(j->st_off + 64, j->ld, wn, 0, cx, W, s.extra_dim);
      g.scatter((uint32_t)j->colsum_off, W, bn, 0, 0, 1, W);
      first = false;
    }
    if (i > 0) {
      mcf_dw_job_t* j = g.job(P, Src{0, fwd.save_h[i], W}, W, W, wn, bn, first);
      g.scatter(j->st_off, j->ld, wn, 0, skip ? cx + s.extra_dim : 0, W, W);
      if (first) g.scatter((uint32_t)j->colsum_off, W, bn, 0, 0, 1, W);
    }
  }
  mcf_dw_job_t* j = g.job(Src{0, fwd.save_h[D], W}, Src{1, bwd.save_ghead, 64}, W, 12, id_final_w(s), -1, false);
  g.scatter(j->st_off, j->ld, id_final_w(s), 0, 0, n_head, W, true);
  const uint32_t hc = g.alloc(12);
  g.scatter(hc, 12, id_final_b(s), 0, 0, 1, n_head);
  out->head_ncols = n_head; out->head_stride = 12; out->head_off = hc;
  if (out->staging_floats < 4) out->staging_floats = 4;
  return g.err;
}

}  // namespace

extern "C" {

int mcf_plan_forward(const mcf_plan_spec_t* spec, mcf_plan_t* out) {
  if (!spec || !out) return MCF_ERR_BAD_ARG;
  const int e = check_spec(*spec);
  if (e) return e;
  return spec->family == 0 ? nerf_forward(*spec, out) : nof_forward(*spec, out);
}

int mcf_plan_backward(const mcf_plan_spec_t* spec, const mcf_plan_t* fwd, mcf_plan_t* out) {
  if (!spec || !fwd || !out) return MCF_ERR_BAD_ARG;
  const int e = check_spec(*spec);
  if (e) return e;
  return spec->family == 0 ? nerf_backward(*spec, *fwd, out) : nof_backward(*spec, *fwd, out);
}

int mcf_plan_gradients(const mcf_plan_spec_t* spec, const mcf_plan_t* fwd, const mcf_plan_t* bwd, mcf_grad_plan_t* out) {
  if (!spec || !fwd || !bwd || !out) return MCF_ERR_BAD_ARG;
  const int e = check_spec(*spec);
  if (e) return e;
  return spec->family == 0 ? nerf_gradients(*spec, *fwd, *bwd, out) : nof_gradients(*spec, *fwd, *bwd, out);
}

}  // extern "C"
