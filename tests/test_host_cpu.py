"""CPU-only checks of the host logic: library exports, struct layouts, plans, image helpers."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import moco_flow_b200 as mf
from moco_flow_b200 import _lib as L
from moco_flow_b200 import plans as P
from tests.helpers import from_images, to_images

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from moco_flow_b200.build import build
    path = build()
    header = open(os.path.join(ROOT, "include", "moco_flow_b200.h")).read()
    declared = set(re.findall(r"\bint\s+(mcf_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    lib = C.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mcf_abi_version() == 1


def test_struct_layouts_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\n'
                   'int main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(mcf_chain_params_t), '
                   'sizeof(mcf_dw_params_t), sizeof(mcf_chunk_t), sizeof(mcf_round_t), sizeof(mcf_pack_t), '
                   'offsetof(mcf_chain_params_t, d_xyz), offsetof(mcf_dw_params_t, n_tiles));return 0;}\n'
                   % os.path.join(ROOT, "include", "moco_flow_b200.h"))
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(L.ChainParams), C.sizeof(L.DwParams), L.CHUNK_DT.itemsize, L.ROUND_DT.itemsize,
            L.PACK_DT.itemsize, L.ChainParams.d_xyz.offset, L.DwParams.n_tiles.offset]
    assert got == want


def test_image_helpers_roundtrip():
    x = torch.randn(300, 100)
    img = to_images(x)
    assert img.shape == (3, 2, 128, 8, 8)
    back = from_images(img, 300, 100)
    assert torch.equal(back, x.to(torch.bfloat16).float())
    # row r chunk c lives at chunk c ^ (r & 7)
    x2 = torch.zeros(128, 64)
    x2[5, 8:16] = 1.0
    assert (to_images(x2)[0, 0, 5, 1 ^ 5, :] == 0x3F80).all()


def test_nerf_plan_shapes():
    pl = P.nerf_forward_plan(8, 256, 63, (4,), 5, sigma_only=False, training=False)
    assert len(pl.chunks) == 72 and pl.wpack_bytes == 72 * 16384
    assert len(pl.rounds) == 10
    assert [int(r["epi"]) for r in pl.rounds] == [0] * 7 + [1, 2, 3]
    assert int(pl.rounds[4]["chunk_end"] - pl.rounds[4]["chunk_begin"]) == 10  # skip layer: x0 + 4 h blocks, 2 halves
    so = P.nerf_forward_plan(8, 256, 63, (4,), 5, sigma_only=True, training=False)
    assert len(so.rounds) == 8 and len(so.chunks) == 60
    tr = P.nerf_forward_plan(8, 256, 63, (4,), 5, sigma_only=False, training=True)
    assert tr.save_tile_bytes == (1 + 8 * 4 + 4 + 2) * 16384   # x0, h1..h8, feat, he (per-ray features: shared images)
    assert tr.mask_tile_words == 8 * 8 * 128 + 4 * 128
    # every chunk consumes columns that exist; first chunk of each accumulator region initialises it
    for r in pl.rounds:
        seen = set()
        for c in pl.chunks[int(r["chunk_begin"]):int(r["chunk_end"])]:
            key = int(c["acc_col"])
            assert bool(c["flags"] & 1) == (key not in seen)
            seen.add(key)


def test_nof_plan_shapes():
    pl = P.nof_forward_plan(4, 128, 33, (2,), 33, True, training=False)
    assert len(pl.rounds) == 5 and pl.n_raybias == 2
    assert [int(r["raybias"]) for r in pl.rounds] == [0, -1, 1, -1, -1]
    assert int(pl.chunks[0]["ksteps"]) == 3  # 33 channels -> K = 48
    assert int(pl.rounds[-1]["epi"]) == L.EPI_NOF_HEAD


def test_unsupported_shapes_fail_loudly():
    with pytest.raises(ValueError):
        P.nerf_forward_plan(8, 192, 63, (4,), 5, False, False)
    with pytest.raises(ValueError):
        P.nerf_forward_plan(8, 256, 93, (4,), 5, False, False)


def test_modules_keep_reference_state_dict_names():
    from oracle import moco_oracle as orc
    n = mf.NeRF(8, 256, 63, [4], "ind", 5)
    assert {k: tuple(v.shape) for k, v in n.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in orc.make_nerf_params(orc.C2F_NERF, 0).items()}
    f = mf.NoF(4, 128, 33, [2], "ind", 33, True)
    assert {k: tuple(v.shape) for k, v in f.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in orc.make_nof_params(orc.C2F_NOF, 0).items()}
    e = mf.Embedding(3, 10)
    assert e.out_channels == 63 and e.weights == [1] * 10
    e.set_weights(0)
    assert e.weights == [0] * 10
    with pytest.raises(AssertionError):
        e.set_weights([1.0, 2.0])
    assert torch.equal(e.freq_bands, 2 ** torch.linspace(0, 9, 10))
    assert torch.equal(mf.Embedding(3, 4, logscale=False).freq_bands, torch.linspace(1, 8, 4))


def test_factories_and_errors():
    m = mf.get_model(dict(type="NoF", D=4, W=128, in_channels_xyz=33, skips=[2], extra_feat_type="ind",
                          extra_feat_dim=33, use_quat=True))
    assert isinstance(m, mf.NoF)
    assert isinstance(mf.get_loss(dict(type="MSE")), mf.MSELoss)
    with pytest.raises(ValueError):
        mf.get_model(dict(type="Foo"))
    with pytest.raises(ValueError):
        mf.get_loss(dict(type="Foo"))
    with pytest.raises(AssertionError):
        mf.NeRF(extra_feat_type="bogus")
    with pytest.raises(AssertionError):
        mf.NoF(extra_feat_type="none")


def test_no_cpu_fallback():
    n = mf.NeRF(8, 256, 63, [4], "ind", 5)
    with pytest.raises(RuntimeError):
        n(torch.zeros(4, 68))
    with pytest.raises(RuntimeError):
        mf.Embedding(3, 2)(torch.zeros(4, 3))


def test_plans_carry_their_program_family():
    """The C ABI picks a kernel instantiation from ``program_kind`` (0 NeRF, 1 NoF): forward and backward plans of a
    module must agree, and the offsets of the struct fields the launcher reads must exist in the ctypes mirror."""
    nf = P.nerf_forward_plan(8, 256, 63, (4,), 5, False, True)
    nb = P.nerf_backward_plan(8, 256, 63, (4,), 5, True, nf)
    of = P.nof_forward_plan(4, 128, 33, (2,), 33, True, True)
    ob = P.nof_backward_plan(4, 128, 33, (2,), 33, True, True, of)
    assert (nf.kind, nb.kind, of.kind, ob.kind) == (0, 0, 1, 1)
    assert all(int(e) < 16 for e in nf.rounds["epi"]) and all(int(e) >= 16 for e in nb.rounds["epi"])
    assert int(of.rounds["epi"][-1]) == L.EPI_NOF_HEAD
    names = [f[0] for f in L.ChainParams._fields_]
    assert names[-8:] == ["cta_pair", "program_kind", "pe_table", "wpack_bytes", "resident", "d_dense", "d_dense_stride",
                          "reserved0"]


def test_bench_reference_arm_contract():
    """``bench.py --impl reference`` (the CPU oracle port; the one place outside tests/ that may execute oracle/) prints
    one JSON line with the contract's keys, on a tiny bounded sample here."""
    import json
    import sys
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0", "--cpu-rays", "16"], stderr=subprocess.DEVNULL, timeout=600)
    line = json.loads(out.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("train rays/s") and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_weights_make_the_self_check_meaningful():
    """bench.py checks its timed computation against the oracle; with nn.Linear's default init the fine-pass density is
    non-positive everywhere (all opacities exactly 0, image = background) and that check would compare constants.  The
    bench's synthetic weights must give semi-transparent-to-opaque rays on the bench's own synthetic rays."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from oracle import moco_oracle as orc
    import torch
    rays, bg, _ = bench.synth_batch(32, seed=1)
    nerf_p, nof_p = bench.synthetic_weights()
    pes = orc.C2F_PE
    with torch.no_grad():
        res = orc.render_rays(rays, bg, [pes["nerf_xyz"], pes["nerf_ind"], None],
                              [orc.NeRFBundle(orc.C2F_NERF, p) for p in nerf_p], [pes["nof_xyz"], pes["nof_ind"]],
                              [orc.NoFBundle(orc.C2F_NOF, p) for p in nof_p], **bench.render_kwargs("render"))
    assert float(res["opacity_fine"].min()) > 0.5
    assert float((res["rgb_fine"] - bg).abs().max()) > 0.05      # not the background
    assert float(res["depth_fine"].std()) > 0.0


def _same_plan(cp, pp, names):
    """C-built plan == plans.py plan: every table entry, size and named offset."""
    from moco_flow_b200 import plans_c
    assert (cp.n_pack, cp.n_chunks, cp.n_rounds) == (len(pp.pack), len(pp.chunks), len(pp.rounds))
    assert [names[cp.tensor_ids[i]] for i in range(cp.n_tensors)] == pp.tensor_names
    for name, dt, n, ref in (("pack", L.PACK_DT, cp.n_pack, pp.pack), ("chunks", L.CHUNK_DT, cp.n_chunks, pp.chunks),
                             ("rounds", L.ROUND_DT, cp.n_rounds, pp.rounds)):
        got = cp.table(name, dt, n)
        assert got.tobytes() == np.ascontiguousarray(ref).tobytes(), name
    assert (cp.wpack_bytes, cp.n_consts, cp.save_tile_bytes, cp.mask_tile_words) == \
        (pp.wpack_bytes, pp.n_consts, pp.save_tile_bytes, pp.mask_tile_words)
    assert (cp.n_raybias, cp.kind, cp.resident, cp.width) == (pp.n_raybias, pp.kind, int(pp.resident), pp.width)
    for key, off in pp.offsets.items():
        kind, what = key.split("_", 1)
        if what[0] == "h" and what[1:].isdigit():
            got = (cp.save_h if kind == "save" else cp.mask_h)[int(what[1:])]
        elif what.startswith("dy") and what[2:].isdigit():
            got = cp.save_dy[int(what[2:])]
        else:
            got = getattr(cp, key)
        assert got == off, key


@pytest.mark.parametrize("nof_kernel", [0, 1, 2])
def test_c_abi_plan_builders_match_python(nof_kernel, monkeypatch):
    """the MLP entries of the C ABI need layer-program tables.  mcf_plan_forward / _backward /
    _gradients build them inside the library; they must equal what plans.py builds, entry by entry."""
    from moco_flow_b200 import plans_c
    from moco_flow_b200.build import build
    build()
    monkeypatch.setattr(P, "NOF_RESIDENT", nof_kernel != 0)
    monkeypatch.setattr(P, "NOF_KERNEL", "ts" if nof_kernel == 2 else "smem")
    cases = [(0, 8, 256, 63, (4,), 5, False), (0, 8, 256, 63, (4,), 27, False), (0, 8, 256, 63, (4,), 0, False),
             (1, 4, 128, 33, (2,), 33, True), (1, 4, 128, 33, (2,), 33, False), (1, 6, 128, 33, (2, 4), 33, True),
             (0, 4, 256, 39, (), 0, False), (0, 8, 256, 63, (2, 5), 5, False), (0, 8, 256, 27, (4,), 5, False),
             (1, 3, 128, 33, (), 33, True), (1, 4, 128, 21, (1,), 8, False)]
    for family, D, W, cx, skips, extra, quat in cases:
        names = plans_c.parameter_names(family, D)
        for training in (False, True):
            for sigma_only in ((False, True) if family == 0 else (False,)):
                s = plans_c.spec(family, D, W, cx, skips, extra, quat, sigma_only, training, nof_kernel=nof_kernel)
                cf = plans_c.forward(s)
                pf = P.nerf_forward_plan(D, W, cx, skips, extra, sigma_only, training) if family == 0 else \
                    P.nof_forward_plan(D, W, cx, skips, extra, quat, training)
                _same_plan(cf, pf, names)
        for need_dx in (False, True):
            s = plans_c.spec(family, D, W, cx, skips, extra, quat, False, True, need_dx, nof_kernel)
            cf = plans_c.forward(s)
            cb = plans_c.backward(s, cf)
            if family == 0:
                pf = P.nerf_forward_plan(D, W, cx, skips, extra, False, True)
                pb = P.nerf_backward_plan(D, W, cx, skips, extra, need_dx, pf)
                mod = mf.NeRF(D, W, cx, list(skips), "ind" if extra else "none", extra)
            else:
                pf = P.nof_forward_plan(D, W, cx, skips, extra, quat, True)
                pb = P.nof_backward_plan(D, W, cx, skips, extra, quat, need_dx, pf)
                mod = mf.NoF(D, W, cx, list(skips), "ind", extra, quat)
            _same_plan(cb, pb, names)
            shapes = {k: tuple(v.shape) for k, v in mod.named_parameters()}
            assert list(shapes) == names
            pg = P.nerf_grad_plan(D, W, cx, skips, extra, shapes, pf, pb) if family == 0 else \
                P.nof_grad_plan(D, W, cx, skips, extra, quat, shapes, pf, pb)
            cg = plans_c.gradients(s, cf, cb)
            assert cg.n_jobs == len(pg.jobs) and cg.n_unpack == len(pg.unpack) and cg.n_params == len(names)
            jobs = np.frombuffer(bytes(cg.jobs), dtype=L.DWJOB_DT)[:cg.n_jobs]
            assert jobs.tobytes() == P.job_table(pg, set(names)).tobytes()
            unp = np.frombuffer(bytes(cg.unpack), dtype=L.UNPACK_DT)[:cg.n_unpack]
            assert unp.tobytes() == np.ascontiguousarray(pg.unpack).tobytes()
            assert [(names[cg.unpack_param[i]], cg.unpack_inner[i]) for i in range(cg.n_unpack)] == pg.unpack_targets
            assert (cg.staging_floats, cg.total_floats) == (pg.staging_floats, pg.total_floats)
            assert (cg.head_ncols, cg.head_stride, cg.head_off) == tuple(pg.head_colsum)
            assert [cg.param_offset[i] for i in range(len(names))] == [pg.param_offsets[n] for n in names]
            for ji, j in enumerate(pg.jobs):
                fed = {names[k] for k in cg.job_params[ji] if k >= 0}
                assert fed == set(j.params), (ji, fed, j.params)


def test_plan_builders_refuse_the_same_shapes():
    """Shapes outside what the fused kernels support: plans.py raises ValueError, the C builders return MCF_ERR_*."""
    from moco_flow_b200 import plans_c
    from moco_flow_b200.build import build
    build()
    for family, D, W, cx, skips, extra in ((0, 8, 192, 63, (4,), 5), (0, 8, 256, 70, (4,), 5), (0, 30, 256, 63, (4,), 5)):
        with pytest.raises(ValueError):
            P.nerf_forward_plan(D, W, cx, skips, extra, False, False)
        with pytest.raises(L.MocoFlowLibraryError):
            plans_c.forward(plans_c.spec(family, D, W, cx, skips, extra, False, False, False))


def test_c_example_compiles(tmp_path):
    """examples/nerf_forward.c drives a NeRF forward through the C ABI alone (plan builder -> pack -> ray bias -> chain
    launch); it must compile against include/moco_flow_b200.h with a plain C compiler."""
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-c", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "nerf_forward.c"), "-o", str(tmp_path / "ex.o")])
