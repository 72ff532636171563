"""Nearest-vertex correspondence sampling."""
import ctypes as C
import os

import pytest
import torch

from oracle import camera_oracle as cam_orc


def _case(V, N, seed):
    gen = torch.Generator().manual_seed(seed)
    verts = torch.randn(V, 3, generator=gen) * 0.4
    query = torch.cat([(torch.rand(N // 2, 3, generator=gen) - 0.5) * 3.0,
                       verts[torch.randint(V, (N - N // 2,), generator=gen)] + torch.randn(N - N // 2, 3, generator=gen) * 0.2])
    trans = torch.eye(4).repeat(V, 1, 1)
    trans[:, :3, :] += torch.randn(V, 3, 4, generator=gen) * 0.1
    return verts, query, trans


def test_oracle_nearest_vertex_properties():
    verts, query, trans = _case(50, 40, 0)
    dist, ind, cano, inside = cam_orc.nearest_vertex(verts, query, trans, 0.2)
    for q in range(query.shape[0]):
        d = (verts - query[q]).norm(dim=1)
        assert abs(float(d.min()) - float(dist[q])) < 1e-6 and int(d.argmin()) == int(ind[q])
    assert torch.equal(inside, dist < 0.2)
    # a query sitting on a vertex maps with that vertex's transform
    d0, i0, c0, _ = cam_orc.nearest_vertex(verts, verts[7:8].clone(), trans, 0.2)
    assert int(i0) == 7 and float(d0) == 0.0
    assert torch.allclose(c0[0], trans[7, :3, :3] @ verts[7] + trans[7, :3, 3], atol=1e-6)


def test_cpu_inputs_raise():
    from moco_flow_b200 import correspondence
    with pytest.raises(RuntimeError):
        correspondence.nearest_vertex(torch.zeros(4, 3), torch.zeros(2, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("V,N", [(6890, 20000), (1000, 257), (3, 1)])
def test_nearest_vertex_kernel(V, N):
    from moco_flow_b200 import correspondence, _lib as L
    dev = torch.device("cuda:0")
    verts, query, trans = _case(V, N, V + N)
    dist_r, ind_r, cano_r, inside_r = cam_orc.nearest_vertex(verts, query, trans, 0.2)
    dist, ind, cano, inside = correspondence.nearest_vertex(verts.to(dev), query.to(dev), trans.to(dev), 0.2)
    torch.cuda.synchronize()
    assert L.device_error_flag() == 0
    dist, ind, cano, inside = dist.cpu(), ind.cpu(), cano.cpu(), inside.cpu()
    assert (dist - dist_r).abs().max().item() <= 1e-6
    same = ind == ind_r
    # a different index is only acceptable on a numerical tie of the two distances
    if not bool(same.all()):
        d_alt = (verts[ind[~same]] - query[~same]).norm(dim=1)
        assert (d_alt - dist_r[~same]).abs().max().item() <= 1e-6
    assert float(same.float().mean()) > 0.999
    assert (cano[same] - cano_r[same]).abs().max().item() <= 1e-5
    near_thr = (dist_r - 0.2).abs() < 1e-6
    assert torch.equal(inside[~near_thr], inside_r[~near_thr])
    ins, outs = correspondence.split_correspondences(query.to(dev), cano.to(dev), inside.to(dev))
    assert ins.shape[0] + outs.shape[0] == N and ins.shape[1] == 6
    # distances only
    d_only, i_only, c_none, _ = correspondence.nearest_vertex(verts.to(dev), query.to(dev))
    assert c_none is None and torch.equal(i_only.cpu(), ind) and torch.equal(d_only.cpu(), dist)


REF_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libknn_cuda_ref.so")


def test_reference_knn_library_is_built_where_the_reference_is():
    """oracle/build_knn_ref.py compiles the reference's own knn.cu (from the wheel inside /root/reference) when the
    reference tree is present; on the GPU box the prebuilt .so travels with the snapshot."""
    from oracle import build_knn_ref
    path = build_knn_ref.build()
    if os.path.exists(build_knn_ref.WHEEL):
        assert path and os.path.exists(path)
        assert hasattr(C.CDLL(path), "knn_ref")


@pytest.mark.gpu
@pytest.mark.parametrize("V,N", [(6890, 20000), (1000, 257), (37, 5)])
def test_nearest_vertex_bit_exact_vs_reference_knn_cuda(V, N):
    """The reference's KNN(k=1, transpose_mode=True) kernel itself (knn_cuda/csrc/cuda/knn.cu, compiled unmodified into
    oracle/_ref) against the product kernel AND the oracle restatement: indices and distances identical, bit for bit,
    at the SMPL size (6890 vertices, datasets/moco_flow_dataset.py:100-122)."""
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libknn_cuda_ref.so not built (python oracle/build_knn_ref.py where /root/reference exists)")
    from moco_flow_b200 import correspondence
    dev = torch.device("cuda:0")
    verts, query, trans = _case(V, N, 3 * V + N)
    # duplicated vertices: exact distance ties, the first index must win in both
    verts[V // 2] = verts[V // 3]
    ref = C.CDLL(REF_LIB)
    ref.knn_ref.restype = C.c_int
    r_t = verts.t().contiguous().to(dev)      # transpose_mode=True: (V,3) -> [dim][V]
    q_t = query.t().contiguous().to(dev)
    scratch = torch.empty(V, N, device=dev)
    ind_ref = torch.empty(1, N, dtype=torch.int64, device=dev)
    rc = ref.knn_ref(C.c_void_p(r_t.data_ptr()), C.c_int(V), C.c_void_p(q_t.data_ptr()), C.c_int(N), C.c_int(3), C.c_int(1),
                     C.c_void_p(scratch.data_ptr()), C.c_void_p(ind_ref.data_ptr()),
                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    dist_ref, ind_ref = scratch[0].cpu(), (ind_ref[0] - 1).cpu()      # knn() returns 1-based indices minus one
    dist, ind, _, _ = correspondence.nearest_vertex(verts.to(dev), query.to(dev), trans.to(dev), 0.2)
    assert torch.equal(ind.cpu(), ind_ref)
    assert torch.equal(dist.cpu(), dist_ref)
    d_o, i_o, _, _ = cam_orc.nearest_vertex(verts, query, trans, 0.2)
    assert torch.equal(i_o, ind_ref) and torch.equal(d_o, dist_ref)
