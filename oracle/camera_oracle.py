"""CPU restatement of the reference's ray generation / canvas scatter.  TEST INFRASTRUCTURE ONLY (see oracle/__init__):
imported by tests/, never by the product path.  Pinned against fixtures produced by the reference's own code:
utils/camera.py (tests/golden/make_golden_camera.py -> tests/golden/camera.npz) and MoCoFlowTrainer.render
(tests/golden/make_golden_canvas.py -> tests/golden/canvas.npz)."""
import numpy as np
import torch


def gen_ray_directions(H, W, focal, camera_c=(0.0, 0.0)):
    """utils/camera.py:29-51: (H, W, 3) camera-frame directions; column index i runs along W, row index j along H;
    no half-pixel offset; BOTH axes are divided by focal[0] (:49)."""
    i = torch.linspace(0, W - 1, W).view(1, W).expand(H, W)
    j = torch.linspace(0, H - 1, H).view(H, 1).expand(H, W)
    f0 = focal[0]
    return torch.stack([(i - camera_c[0]) / f0, -(j - camera_c[1]) / f0, -torch.ones_like(i)], -1)


def gen_rays(directions, c2w):
    """utils/camera.py:53-82."""
    if c2w is None:
        rays_d = directions / torch.norm(directions, dim=-1, keepdim=True)
        rays_o = torch.zeros_like(directions)
    else:
        rays_d = directions @ c2w[:, :3].T
        rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
        rays_o = c2w[:, 3].expand(rays_d.shape)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)


def make_rays(H, W, K, c2w, aabb_verts, idx):
    """Camera.__init__ + Camera.make_rays, utils/camera.py:98-110,134-148."""
    focal = [K[0][0], K[1][1]]
    center = [K[0][2], K[1][2]]
    d = np.sqrt(np.sum((aabb_verts - c2w[:3, 3]) ** 2, axis=-1))
    near, far = min(d), max(d)
    rays_o, rays_d = gen_rays(gen_ray_directions(H, W, focal, center), torch.from_numpy(c2w[:3, :4]).float())
    one = torch.ones_like(rays_o[:, :1])
    return torch.cat([rays_o, rays_d, near * one, far * one, idx * one], 1)


def canvas_scatter(background, rays_msk, rgb, depth, opacity):
    """trainer/trainer_moco_flow.py:247-262 on CPU tensors; ``rays_msk`` bool (P,), the rest per masked ray."""
    P = background.shape[0]
    msk = np.where(rays_msk)
    img_raw = torch.zeros(P, 3)
    depth_raw = torch.ones(P) * 10
    op = opacity.numpy()
    foreground_idx = np.where(op > 0)
    foreground_mask = np.zeros(P, dtype=np.float64)
    foreground_mask[msk] = op
    img_raw[foreground_mask > 0] = rgb[foreground_idx]
    depth_raw[msk] = 8
    depth_raw[foreground_mask > 0] = depth[foreground_idx]
    img_raw[foreground_mask == 0] = background[foreground_mask == 0]
    return img_raw, depth_raw


def _fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64; one rounding of the sum to float64
    and one to float32 (a double rounding differs from a true fma only on exact float64 half-way cases)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def nearest_vertex(verts, query, trans, thickness):
    """datasets/moco_flow_dataset.py:121-130.  ``knn_cuda.KNN(k=1, transpose_mode=True)`` restated from its source,
    which ships in the reference tree (docker/KNN_CUDA-0.2-py3-none-any.whl: knn_cuda/csrc/cuda/knn.cu):
      cuComputeDistanceGlobal  tmp = ref - query per coordinate, ``ssd += tmp*tmp`` over x, y, z in that order (nvcc's
                               default -fmad=true contracts each to an FMA; the first one adds to 0);
      cuInsertionSort (k = 1)  keeps the first strict minimum over the reference index;
      cuParallelSqrt           sqrt of the kept squared distance;  ``knn()`` returns the index 0-based.
    PINNED: tests/test_correspondence.py compares this restatement and the CUDA kernel bit for bit with the reference's
    own knn.cu, compiled by oracle/build_knn_ref.py into oracle/_ref/libknn_cuda_ref.so."""
    v, q = verts.numpy().astype(np.float32), query.numpy().astype(np.float32)
    best = np.full(q.shape[0], np.inf, np.float32)
    ind = np.zeros(q.shape[0], np.int64)
    step = 256
    for b in range(0, v.shape[0], step):
        d = v[None, b:b + step, :] - q[:, None, :]                      # ref - query
        ssd = (d[..., 0] * d[..., 0]).astype(np.float32)                 # fma(dx, dx, 0)
        ssd = _fma32(d[..., 1], d[..., 1], ssd)
        ssd = _fma32(d[..., 2], d[..., 2], ssd)
        loc = ssd.argmin(axis=1)                                         # first minimum inside the block
        val = ssd[np.arange(q.shape[0]), loc]
        better = val < best                                              # strict: earlier blocks win ties
        best = np.where(better, val, best)
        ind = np.where(better, loc + b, ind)
    dist = torch.from_numpy(np.sqrt(best))
    ind = torch.from_numpy(ind)
    homo = torch.cat([query, torch.ones(query.shape[0], 1)], dim=-1)
    cano = (trans[ind] @ homo.unsqueeze(-1))[:, :3, 0]
    return dist, ind, cano, dist < thickness
