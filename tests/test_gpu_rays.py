"""-m gpu: camera rays and per-object ray assembly through the C ABI, against the
reference-generated golden fixtures (tests/golden/rays_*.npz) and the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import onerf_oracle as O
from tests import cases

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


class Box:          # the attributes of BBoxRayHelper that the ray kernels read
    def __init__(self, inp):
        self.pose_avg, self.axis_align_mat, self.bbox_bounds = inp["pose_avg"], inp["axis_align_mat"], inp["bbox_bounds"]
        self.scale_factor = inp["scale_factor"]


@pytest.mark.parametrize("name", list(cases.CAMERA_CASES))
def test_camera_rays_match_reference_golden(golden, name):
    from object_nerf_b200 import ray_utils
    inp = cases.build_camera_case(cases.CAMERA_CASES[name])
    gold = golden("rays_" + name)
    directions = ray_utils.get_ray_directions(inp["H"], inp["W"], inp["focal"], device=DEV)
    assert directions.is_cuda and tuple(directions.shape) == (inp["H"], inp["W"], 3)
    assert torch.equal(directions.cpu(), gold["directions"])           # fp32 sub / div: bit-exact
    rays_o, rays_d = ray_utils.get_rays(directions, inp["c2w"])
    assert torch.equal(rays_o.cpu(), gold["rays_o"])
    # the 3-term dot product order of the reference's CPU sgemm is unspecified: 2 ulp of a unit vector
    assert (rays_d.cpu() - gold["rays_d"]).abs().max().item() <= 2.5e-7
    assert (rays_d.norm(dim=-1) - 1).abs().max().item() <= 2e-7
    # fused: pixel grid + pose -> (N,8) scene rays
    rays = ray_utils.camera_rays(inp["H"], inp["W"], inp["focal"], inp["c2w"], near=0.3, far=6.0, scale_factor=2.0, device=DEV)
    assert torch.equal(rays[:, :3], rays_o) and torch.equal(rays[:, 3:6], rays_d)
    want = O.generate_rays(0, gold["rays_o"], gold["rays_d"], 0.3, 6.0, 2.0)
    assert torch.equal(rays[:, 6:].cpu(), want[:, 6:])


@pytest.mark.parametrize("name", list(cases.BBOX_CASES))
def test_ray_bbox_intersections_match_reference_golden(golden, name):
    from object_nerf_b200 import ray_utils
    inp = cases.build_bbox_case(cases.BBOX_CASES[name])
    gold = golden("rays_" + name)
    box = Box(inp)
    o, d = inp["rays_o"].to(DEV), inp["rays_d"].to(DEV)
    mask, near, far = ray_utils.get_ray_bbox_intersections(box, o, d, inp["scale_factor"], inp["bbox_enlarge"])
    gmask = gold["mask"].bool()
    assert mask.dtype == torch.bool and tuple(near.shape) == (o.shape[0], 1)
    assert torch.equal(mask.cpu(), gmask)                                # the hit mask is index-like: exact
    assert gmask.any() and (~gmask).any()
    # float64 slab test rounded once to fp32: the only freedom is the association of the 3-term float64 dot products
    for got, want, k in ((near, gold["near"], "near"), (far, gold["far"], "far")):
        err = (got.cpu() - want).abs()
        assert (err <= 1e-6 * want.abs() + 1e-9).all(), (k, err.max().item())
        assert (got.cpu()[~gmask] == 0).all()
    # generate_rays: the (N,8) object rays the renderer feeds to render_rays_multi
    rays = ray_utils.generate_rays(4, o, d, inp["near"], inp["far"], inp["scale_factor"], box=box, bbox_enlarge=inp["bbox_enlarge"])
    want = O.generate_rays(4, inp["rays_o"], inp["rays_d"], inp["near"], inp["far"], inp["scale_factor"],
                           box=dict(pose_avg=inp["pose_avg"], axis_align_mat=inp["axis_align_mat"], bbox_bounds=inp["bbox_bounds"]),
                           bbox_enlarge=inp["bbox_enlarge"])
    assert torch.equal(rays[:, :6].cpu(), want[:, :6])
    assert ((rays[:, 6:].cpu() - want[:, 6:]).abs() <= 1e-6 * want[:, 6:].abs() + 1e-9).all()
    # scene rays (obj_id 0): constants
    rays0 = ray_utils.generate_rays(0, o, d, inp["near"], inp["far"], inp["scale_factor"])
    assert torch.equal(rays0.cpu(), O.generate_rays(0, inp["rays_o"], inp["rays_d"], inp["near"], inp["far"], inp["scale_factor"]))


def test_fused_camera_rays_with_box_equal_the_three_steps():
    """onerf_camera_rays == get_ray_directions -> get_rays -> generate_rays, bit for bit, at frame size."""
    from object_nerf_b200 import ray_utils
    cam = cases.build_camera_case(dict(H=480, W=640, fovx_deg=70.0, seed=321))
    inp = cases.build_bbox_case(cases.BBOX_CASES["bbox_basic"])
    box = Box(inp)
    d = ray_utils.get_ray_directions(cam["H"], cam["W"], cam["focal"], device=DEV)
    ro, rd = ray_utils.get_rays(d, cam["c2w"])
    steps, m1 = ray_utils.generate_rays(7, ro, rd, 0.3, 6.0, 2.0, box=box, bbox_enlarge=0.05, return_mask=True)
    fused, m2 = ray_utils.camera_rays(cam["H"], cam["W"], cam["focal"], cam["c2w"], 0.3, 6.0, 2.0, box=box, bbox_enlarge=0.05,
                                      device=DEV, return_mask=True)
    assert torch.equal(fused, steps) and torch.equal(m1, m2)
    assert tuple(fused.shape) == (480 * 640, 8)
    # size-independent properties: unit directions, near <= far, misses are exactly (0, 0)
    assert (fused[:, 3:6].norm(dim=-1) - 1).abs().max().item() <= 2e-7
    assert (fused[:, 6] <= fused[:, 7]).all() and (fused[~m2][:, 6:] == 0).all()
