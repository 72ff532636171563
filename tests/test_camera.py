"""Ray generation / canvas scatter: oracle vs the reference's fixtures on CPU, kernels vs both on GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import camera_oracle as cam_orc


@pytest.fixture(scope="module")
def golden(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "camera.npz")))


@pytest.mark.parametrize("name", ["a", "b"])
def test_camera_oracle_matches_reference_fixture(golden, name):
    g = golden
    H, W = int(g[f"{name}_H"]), int(g[f"{name}_W"])
    K = g[f"{name}_K"]
    dirs = cam_orc.gen_ray_directions(H, W, [K[0][0], K[1][1]], [K[0][2], K[1][2]])
    assert np.array_equal(dirs.numpy(), g[f"{name}_dirs"])
    rays = cam_orc.make_rays(H, W, K, g[f"{name}_c2w"], g[f"{name}_verts"], float(g[f"{name}_idx"]))
    assert np.array_equal(rays.numpy(), g[f"{name}_rays"])
    _, d_cam = cam_orc.gen_rays(dirs, None)
    assert np.array_equal(d_cam.numpy(), g[f"{name}_rays_d_cam"])


@pytest.fixture(scope="module")
def canvas_golden(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "canvas.npz")))


@pytest.mark.parametrize("name", ["fine", "coarse"])
def test_canvas_oracle_matches_reference_fixture(canvas_golden, name):
    """The image / depth assembly of the reference's own MoCoFlowTrainer.render (trainer/trainer_moco_flow.py:249-263,
    run by tests/golden/make_golden_canvas.py) against the oracle restatement: bit for bit."""
    g = canvas_golden
    img, dep = cam_orc.canvas_scatter(torch.from_numpy(g[f"{name}_background"]), g[f"{name}_rays_msk"],
                                      torch.from_numpy(g[f"{name}_rgb"]), torch.from_numpy(g[f"{name}_depth"]),
                                      torch.from_numpy(g[f"{name}_opacity"]))
    assert np.array_equal(img.numpy(), g[f"{name}_img"]) and np.array_equal(dep.numpy(), g[f"{name}_depth_img"])
    assert (g[f"{name}_opacity"] == 0).any() and (g[f"{name}_depth_img"] == 8).any() and (g[f"{name}_depth_img"] == 10).any()


def test_canvas_oracle_semantics():
    P = 10
    bg = torch.arange(P * 3, dtype=torch.float32).view(P, 3)
    msk = np.zeros(P, dtype=bool)
    msk[[1, 4, 7]] = True
    rgb = torch.tensor([[.1, .2, .3], [.4, .5, .6], [.7, .8, .9]])
    depth = torch.tensor([2.5, 3.5, 4.5])
    opacity = torch.tensor([0.9, 0.0, 0.2])
    img, dep = cam_orc.canvas_scatter(bg, msk, rgb, depth, opacity)
    assert torch.equal(img[1], rgb[0]) and torch.equal(img[7], rgb[2]) and torch.equal(img[4], bg[4])
    assert dep.tolist() == [10, 2.5, 10, 10, 8, 10, 10, 4.5, 10, 10]


def test_camera_cpu_inputs_raise():
    from moco_flow_b200 import camera
    with pytest.raises(RuntimeError):
        camera.make_rays(4, 4, 5.0, [2, 2], np.eye(4)[:3], 1.0, 2.0, 0.0, device="cpu")
    with pytest.raises(RuntimeError):
        camera.scatter_canvas(torch.zeros(4, 3), None, torch.zeros(4, 3), torch.zeros(4), torch.zeros(4))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b"])
def test_make_rays_kernel(golden, name):
    """All pixels, a masked subset (the valid-ray gather) and the camera-frame variant vs the reference fixture.
    fp32: origins / near / far / idx exact, directions <= 2 ulp (the reference's matmul may fuse multiply-adds)."""
    from moco_flow_b200 import camera, _lib as L
    g = golden
    dev = torch.device("cuda:0")
    H, W = int(g[f"{name}_H"]), int(g[f"{name}_W"])
    K, c2w, verts, idx = g[f"{name}_K"], g[f"{name}_c2w"], g[f"{name}_verts"], float(g[f"{name}_idx"])
    cam = camera.Camera((H, W), K, device=dev)
    cam.c2w = c2w
    rays = cam.make_rays(verts, idx).cpu().numpy()
    ref = g[f"{name}_rays"]
    assert rays.shape == ref.shape
    assert np.array_equal(rays[:, :3], ref[:, :3]) and np.array_equal(rays[:, 6:], ref[:, 6:])
    assert np.abs(rays[:, 3:6] - ref[:, 3:6]).max() <= 2.5e-7
    pick = torch.tensor(sorted(np.random.default_rng(1).choice(H * W, 37, replace=False).tolist()), device=dev)
    sub = cam.make_rays(verts, idx, pixel_index=pick).cpu().numpy()
    assert np.array_equal(sub, rays[pick.cpu().numpy()])
    near, far = camera.near_far_from_aabb(verts, c2w)
    cam_frame = camera.make_rays(H, W, cam.focal[0], cam.center, None, near, far, idx, device=dev).cpu().numpy()
    assert np.abs(cam_frame[:, 3:6] - g[f"{name}_rays_d_cam"]).max() <= 2.5e-7 and not cam_frame[:, :3].any()
    assert L.device_error_flag() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["fine", "coarse"])
def test_canvas_scatter_kernel_vs_reference_fixture(canvas_golden, name):
    """The device-side assembly against the output of the reference's own render(): bit for bit."""
    from moco_flow_b200 import camera
    g = canvas_golden
    dev = torch.device("cuda:0")
    pix = torch.from_numpy(np.where(g[f"{name}_rays_msk"])[0]).to(dev)
    bg, rgb, depth, opacity = (torch.from_numpy(g[f"{name}_{k}"]).to(dev) for k in ("background", "rgb", "depth", "opacity"))
    img, dep = camera.scatter_canvas(bg, pix, rgb, depth, opacity)
    assert np.array_equal(img.cpu().numpy(), g[f"{name}_img"])
    assert np.array_equal(dep.cpu().numpy(), g[f"{name}_depth_img"])


@pytest.mark.gpu
def test_canvas_scatter_kernel():
    from moco_flow_b200 import camera
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(4)
    P, n = 5000, 1777
    bg = torch.rand(P, 3, generator=gen)
    msk = np.zeros(P, dtype=bool)
    msk[np.random.default_rng(2).choice(P, n, replace=False)] = True
    rgb, depth = torch.rand(n, 3, generator=gen), torch.rand(n, generator=gen) + 2
    opacity = torch.rand(n, generator=gen)
    opacity[opacity < 0.3] = 0.0
    img_ref, dep_ref = cam_orc.canvas_scatter(bg, msk, rgb, depth, opacity)
    pix = torch.from_numpy(np.where(msk)[0]).to(dev)
    img, dep = camera.scatter_canvas(bg.to(dev), pix, rgb.to(dev), depth.to(dev), opacity.to(dev))
    assert torch.equal(img.cpu(), img_ref) and torch.equal(dep.cpu(), dep_ref)
    # no mask: every pixel rendered
    img2, dep2 = camera.scatter_canvas(bg[:n].to(dev), None, rgb.to(dev), depth.to(dev), opacity.to(dev))
    ref2 = cam_orc.canvas_scatter(bg[:n], np.ones(n, dtype=bool), rgb, depth, opacity)
    assert torch.equal(img2.cpu(), ref2[0]) and torch.equal(dep2.cpu(), ref2[1])
