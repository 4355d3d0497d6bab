"""Shared test-case definitions: the same dicts drive tools/make_golden.py (real reference, build
container), the CPU oracle tests and the GPU parity tests.  Inputs are regenerated from seeds."""
from __future__ import annotations

import numpy as np
import torch

from . import synth

GRID_KW = dict(seed=5, shape=(42, 42, 22), occupancy=0.6, voxel_size=0.05)

_BASE = dict(n_rays=48, n_samples=64, n_importance=64, use_voxel=True, use_disp=False, perturb=0.0,
             noise_std=0.0, white_back=False, forward_instance=True, frustum_bound_th=0.0,
             rays_in_bbox=False, is_eval=True, pass_through=False, sigma_gain=8.0, sigma_bias=1.0,
             seed=100)

RENDER_CASES = {
    # BASELINE.json configs[0]: single 64-ray chunk, 64 coarse samples, scene branch only
    "cfg1_plain": dict(_BASE, n_rays=64, n_importance=0, use_voxel=False, forward_instance=False, seed=101),
    "cfg1_voxel": dict(_BASE, n_rays=64, n_importance=0, forward_instance=False, seed=102),
    # configs[1]-like: 64 coarse + 64 importance (128-sample fine pass), two-branch, eval
    "eval_voxel": dict(_BASE, seed=103),
    "eval_plain": dict(_BASE, use_voxel=False, seed=104),
    "eval_in_bbox": dict(_BASE, rays_in_bbox=True, seed=105),
    "eval_white_disp": dict(_BASE, white_back=True, use_disp=True, n_rays=33, seed=106),
    # configs[2]-like training-mode forward: jitter + sigma noise + occlusion mask + pass-through
    "train_voxel": dict(_BASE, perturb=1.0, noise_std=1.0, is_eval=False, frustum_bound_th=0.025,
                        pass_through=True, seed=107),
    "train_nomask": dict(_BASE, perturb=1.0, noise_std=1.0, is_eval=False, frustum_bound_th=-1.0 / 16,
                         seed=108, n_rays=17),
    # ragged sizes / odd sample counts
    "odd_sizes": dict(_BASE, n_rays=5, n_samples=40, n_importance=24, seed=109),
    "one_ray": dict(_BASE, n_rays=1, seed=110),
}

MULTI_CASES = {
    # configs[4]: scene + duplicated object (ids [0,4,4]), removed-object boxes on the scene branch
    "edit_dup": dict(n_rays=40, n_samples=64, n_importance=64, obj_ids=[0, 4, 4], white_back=False,
                     boxes=True, seed=200),
    "edit_scene_only": dict(n_rays=16, n_samples=64, n_importance=64, obj_ids=[0], white_back=True,
                            boxes=False, seed=201),
    "edit_two_objs": dict(n_rays=24, n_samples=32, n_importance=32, obj_ids=[4, 6], white_back=False,
                          boxes=False, seed=202),
}


def build_render_case(c):
    s = c["seed"]
    w = {"coarse": synth.make_weights(s, c["use_voxel"], c["sigma_gain"], c["sigma_bias"])}
    if c["n_importance"] > 0:
        w["fine"] = synth.make_weights(s + 1000, c["use_voxel"], c["sigma_gain"], c["sigma_bias"])
    n = c["n_rays"]
    rays = synth.random_rays(s + 1, n)
    code_table = synth.make_codes(s + 2)
    rng = np.random.default_rng(s + 3)
    ids = rng.choice([4, 6], size=n)
    codes = code_table[torch.from_numpy(ids)]
    ptm = torch.from_numpy(rng.random((n, 1)) < 0.5) if c["pass_through"] else None
    return {
        "weights": w,
        "grid": synth.make_grid(**GRID_KW) if c["use_voxel"] else None,
        "rays": rays,
        "codes": codes,
        "pass_through_mask": ptm,
        "rand": synth.random_buffers(s + 4, n, c["n_samples"], c["n_importance"]),
    }


def build_multi_case(c):
    s = c["seed"]
    n = c["n_rays"]
    w = {"coarse": synth.make_weights(s, True, 8.0, 1.0), "fine": synth.make_weights(s + 1000, True, 8.0, 1.0)}
    rng = np.random.default_rng(s + 5)
    rays_list = []
    for k, iid in enumerate(c["obj_ids"]):
        rays = synth.random_rays(s + 10 + k, n) if k == 0 else rays_list[0].clone()
        if iid > 0:
            # object ray sets: per-ray near/far from a bbox hit; misses get near = far = 0
            # (render_tools/editable_renderer.py:153-181)
            near = torch.from_numpy(rng.uniform(0.4, 1.2, size=n).astype(np.float32))
            far = near + torch.from_numpy(rng.uniform(0.2, 0.9, size=n).astype(np.float32))
            miss = torch.from_numpy(rng.random(n) < 0.3)
            near[miss] = 0
            far[miss] = 0
            rays = rays.clone()
            rays[:, 0:3] += torch.from_numpy(rng.normal(0, 0.05, size=(1, 3)).astype(np.float32))
            rays[:, 6], rays[:, 7] = near, far
        rays_list.append(rays)
    boxes = []
    if c["boxes"]:
        for b in range(2):
            ang = 0.3 + 0.5 * b
            A = np.eye(4)
            A[:3, :3] = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
            A[:3, 3] = np.array([0.1 * b, -0.2, 0.05])
            P = np.eye(4)
            P[:3, 3] = np.array([0.3, -0.1 * b, 0.0])
            lo = np.array([-0.5, -0.4, -0.3]) + 0.1 * b
            boxes.append(dict(scale_factor=2.0, pose_avg=P, axis_align_mat=A, bbox_bounds=np.array([lo, lo + 0.9])))
    return {"weights": w, "grid": synth.make_grid(**GRID_KW), "rays_list": rays_list,
            "code_table": synth.make_codes(s + 2), "boxes": boxes}


def box_affine(b):
    """Fold BBoxRayHelper.transform_xyz_to_bbox_coordinates (utils/bbox_utils.py:119-130: unscale,
    pose_avg, axis_align) into p_box = A p + t; bounds from check_xyz_in_bounds with bbox_enlarge = 0."""
    sf = b["scale_factor"]
    P, Ax = b["pose_avg"], b["axis_align_mat"]
    M = Ax[:3, :3] @ P[:3, :3]
    A = M * sf
    t = Ax[:3, :3] @ P[:3, 3] + Ax[:3, 3]
    return (torch.from_numpy(A).float(), torch.from_numpy(t).float(),
            torch.from_numpy(b["bbox_bounds"][0]).float(), torch.from_numpy(b["bbox_bounds"][1]).float())


def stage_inputs():
    rng = np.random.default_rng(7)
    f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    n = 96
    g = synth.make_grid(**GRID_KW)
    ext = (g["shape"].numpy() * float(g["voxel_size"]))
    xyz = rng.uniform(-0.62, 0.62, size=(160, 3)) * ext  # inside and beyond the volume
    wts = rng.random((n, 62)) ** 4
    wts[:4] = 0.0            # all-zero rows (denominator guard)
    wts[4:8, 10:] = 0.0      # empty tails
    near = rng.uniform(0.1, 0.3, size=(n, 1))
    bins = near + np.sort(rng.random((n, 63)), -1) * 2.5
    return {
        "posenc_x": f(rng.uniform(-3, 3, size=(64, 3))),
        "voxel_xyz": f(xyz),
        "emb_xyz_v": f(rng.uniform(-1, 1, size=(n, 271))),
        "emb_xyz_p": f(rng.uniform(-1, 1, size=(n, 63))),
        "emb_dir": f(rng.uniform(-1, 1, size=(n, 27))),
        "obj_voxel": f(rng.uniform(-1, 1, size=(n, 104))),
        "obj_code": f(rng.standard_normal((n, 64))),
        "pdf_bins": f(bins),
        "pdf_weights": f(wts),
        "pdf_u": f(rng.random((n, 64))),
    }


# ---------------------------------------------------------------------------------------------
# training-step (gradient) case: SURVEY.md §8d config 3 in miniature
# ---------------------------------------------------------------------------------------------
GRAD_CASE = dict(_BASE, n_rays=40, perturb=1.0, noise_std=1.0, is_eval=False, frustum_bound_th=0.025,
                 pass_through=True, seed=300, sigma_gain=4.0, sigma_bias=0.5)
LOSS_CONF = dict(color_loss_weight=1.0, depth_loss_weight=0.1, opacity_loss_weight=100.0,
                 instance_color_loss_weight=1.0, instance_depth_loss_weight=0.1)   # default_conf.yml:61-66 + scannet override
GRAD_SAMPLES = 256   # entries sampled per parameter tensor in the fixture


def build_grad_case(n_rays=None):
    """n_rays=None: the fixture's case (tools/make_golden.py); other sizes reuse its seeds on more rays."""
    c = GRAD_CASE if n_rays is None else dict(GRAD_CASE, n_rays=n_rays)
    inp = build_render_case(c)
    n = c["n_rays"]
    rng = np.random.default_rng(c["seed"] + 9)
    ids = rng.choice([4, 6], size=n)
    inp["instance_ids"] = torch.from_numpy(ids).view(n, 1)
    inp["code_table"] = synth.make_codes(c["seed"] + 2)
    inp["batch"] = {
        "rgbs": torch.from_numpy(rng.random((n, 3)).astype(np.float32)),
        "depths": torch.from_numpy(rng.uniform(0.3, 2.5, size=n).astype(np.float32)),
        "valid_mask": torch.from_numpy(rng.random(n) < 0.9),
        "instance_mask": torch.from_numpy(rng.random(n) < 0.5),
        "instance_mask_weight": torch.from_numpy(np.where(rng.random(n) < 0.5, 1.0, 0.05).astype(np.float32)),
    }
    return inp


def total_loss(out, batch, conf=LOSS_CONF):
    """The reference's TotalLoss (models/losses.py:5-135) restated for tests: masked MSEs over coarse + fine maps."""
    vm = batch["valid_mask"].view(-1)
    im = batch["instance_mask"].view(-1)
    imw = batch["instance_mask_weight"].view(-1)
    tgt_rgb, tgt_d = batch["rgbs"].view(-1, 3), batch["depths"].view(-1)
    loss = 0.0
    for typ in ("coarse", "fine"):
        if f"rgb_{typ}" not in out:
            continue
        loss = loss + conf["color_loss_weight"] * ((out[f"rgb_{typ}"][vm] - tgt_rgb[vm]) ** 2).mean()
        dm = vm & (tgt_d > 0)
        loss = loss + conf["depth_loss_weight"] * ((out[f"depth_{typ}"][dm] - tgt_d[dm]) ** 2).mean()
        loss = loss + conf["opacity_loss_weight"] * (
            ((out[f"opacity_instance_{typ}"][vm].clamp(0, 1) - im[vm].float()) ** 2) * imw[vm]).mean()
        m2 = vm & im
        loss = loss + conf["instance_color_loss_weight"] * (
            ((out[f"rgb_instance_{typ}"][m2] - tgt_rgb[m2]) ** 2) * imw[m2][:, None]).mean()
        m3 = dm & im
        loss = loss + conf["instance_depth_loss_weight"] * (
            ((out[f"depth_instance_{typ}"][m3] - tgt_d[m3]) ** 2) * imw[m3]).mean()
    return loss


def sample_indices(name, numel):
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    return torch.from_numpy(rng.integers(0, numel, size=min(GRAD_SAMPLES, numel)))


# ------------------------------------------------------------------------------------------------
# camera rays / per-object ray assembly
# ------------------------------------------------------------------------------------------------
CAMERA_CASES = {
    "cam_small": dict(H=24, W=32, fovx_deg=70.0, seed=301),
    "cam_odd": dict(H=37, W=53, fovx_deg=55.0, seed=302),
}


def build_camera_case(c):
    """focal as editable_renderer.py:190, a random rigid camera-to-world pose (3,4) fp32."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(c["seed"]))
    focal = (c["W"] / 2) / np.tan((c["fovx_deg"] / 2) / (180 / np.pi))
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    t = rng.uniform(-1.0, 1.0, size=3)
    c2w = torch.from_numpy(np.concatenate([q, t[:, None]], 1).astype(np.float32))
    return dict(H=c["H"], W=c["W"], focal=float(focal), c2w=c2w)


BBOX_CASES = {
    "bbox_basic": dict(n_rays=4096, enlarge=0.0, scale=2.0, seed=311),
    "bbox_enlarged": dict(n_rays=2048, enlarge=0.07, scale=16.0, seed=312),
}


def build_bbox_case(c):
    """Rays around a box given in a rotated / translated frame: most aimed at the box, some missing it, some starting
    INSIDE it (a miss by the reference's rule), some with exactly-zero direction components (the 1e-14 rule)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(c["seed"]))
    n, s = c["n_rays"], c["scale"]

    def rigid(rotate):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        T = np.eye(4)
        if rotate:
            T[:3, :3] = q
        T[:3, 3] = rng.uniform(-0.5, 0.5, size=3)
        return T

    # pose_avg is a pure translation, as in the datasets: the reference rotates the DIRECTION by the axis-alignment
    # matrix only (utils/bbox_utils.py:116), which is consistent only then
    pose_avg, axis = rigid(False), rigid(True)
    lo = np.array([-0.6, -0.4, -0.3]) + rng.uniform(-0.05, 0.05, 3)
    hi = np.array([0.5, 0.7, 0.4]) + rng.uniform(-0.05, 0.05, 3)
    bounds = np.stack([lo, hi])
    # choose world-space (NeRF scale) origins / targets by mapping box-frame points back
    Tinv = np.linalg.inv(axis @ pose_avg)
    def to_world(pb):
        return ((Tinv[:3, :3] @ pb.T).T + Tinv[:3, 3]) / s
    centre, half = (lo + hi) / 2, (hi - lo) / 2
    tgt_box = centre + rng.uniform(-1.4, 1.4, size=(n, 3)) * half          # ~50 % inside the box footprint
    org_box = centre + rng.normal(size=(n, 3)) * 3.0
    inside = rng.uniform(size=n) < 0.1
    org_box[inside] = centre + rng.uniform(-0.9, 0.9, size=(int(inside.sum()), 3)) * half
    o = to_world(org_box).astype(np.float32)
    d = to_world(tgt_box) - to_world(org_box)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    zero = rng.uniform(size=n) < 0.05
    d[zero, rng.integers(0, 3, size=int(zero.sum()))] = 0.0
    return dict(rays_o=torch.from_numpy(o), rays_d=torch.from_numpy(d), pose_avg=pose_avg, axis_align_mat=axis,
                bbox_bounds=bounds, scale_factor=float(s), bbox_enlarge=float(c["enlarge"]), near=0.3, far=6.0)


# ------------------------------------------------------------------------------------------------
# training loss
# ------------------------------------------------------------------------------------------------
LOSS_CASES = {
    "loss_train": dict(n=2048, fine=True, p_valid=0.9, p_inst=0.5, p_depth=0.8, seed=401),
    "loss_coarse_only": dict(n=300, fine=False, p_valid=0.7, p_inst=0.3, p_depth=0.5, seed=402),
    "loss_no_instance": dict(n=257, fine=True, p_valid=0.8, p_inst=0.0, p_depth=0.6, seed=403),   # instance terms skipped
    "loss_no_depth": dict(n=64, fine=True, p_valid=1.0, p_inst=0.5, p_depth=0.0, seed=404),       # depth terms skipped
}
LOSS_MAP_KEYS = ("rgb", "depth", "opacity_instance", "rgb_instance", "depth_instance")


def build_loss_case(c):
    """Random rendered maps (opacity outside [0, 1] for some rays: the clamp) and a random batch
    (datasets/generic_dataset.py keys consumed by models/losses.py)."""
    rng = np.random.Generator(np.random.PCG64(c["seed"]))
    n = c["n"]
    f = lambda *shape: torch.from_numpy(rng.uniform(0, 1, size=shape).astype(np.float32))
    maps = {}
    for typ in (("coarse", "fine") if c["fine"] else ("coarse",)):
        maps[f"rgb_{typ}"], maps[f"rgb_instance_{typ}"] = f(n, 3), f(n, 3)
        maps[f"depth_{typ}"], maps[f"depth_instance_{typ}"] = f(n) * 3, f(n) * 3
        maps[f"opacity_instance_{typ}"] = f(n) * 1.4 - 0.2
    depths = f(n) * 3
    depths[torch.from_numpy(rng.uniform(size=n) >= c["p_depth"])] = 0.0
    im = torch.from_numpy(rng.uniform(size=n) < c["p_inst"])
    batch = dict(rgbs=f(n, 3), depths=depths, valid_mask=torch.from_numpy(rng.uniform(size=n) < c["p_valid"]),
                 instance_mask=im, instance_mask_weight=torch.where(im, torch.tensor(1.0), torch.tensor(0.05)))
    return maps, batch


# ------------------------------------------------------------------------------------------------
# voxel grid maintenance
# ------------------------------------------------------------------------------------------------
MAINT_CASE = dict(n_points=44, seed=501, max_voxels=4096, max_alpha_th=0.1236, sigma_gain=6.0, sigma_bias=0.0,
                  extra=dict(pcd_path="maint", scene_center=[0.5, 0.5, 0.0], scale_factor=2.0, voxel_size=0.4, neighbor_marks=1))


def build_maint_case():
    """A tiny cloud (a few dozen occupied voxels: the reference prunes with 16^3 samples per voxel), a random feature
    table and the fine model's weights."""
    c = MAINT_CASE
    rng = np.random.Generator(np.random.PCG64(c["seed"]))
    pts = rng.uniform([0, 0, -0.6], [2.4, 2.0, 0.6], size=(c["n_points"], 3))
    table = torch.from_numpy(rng.normal(0, 0.35, size=(c["max_voxels"], 24)).astype(np.float32))
    weights = synth.make_weights(c["seed"] + 1, use_voxel=True, sigma_gain=c["sigma_gain"], sigma_bias=c["sigma_bias"])
    return dict(points=pts, table=table, weights=weights)


def maint_rand(n_chunks, seed=MAINT_CASE["seed"] + 7):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(32 * 16 ** 3, 3, generator=g) for _ in range(n_chunks)]
