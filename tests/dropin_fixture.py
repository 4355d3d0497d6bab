"""A synthetic ScanNet-style scene on disk (point cloud, axis alignment, instance boxes, camera poses, training config
snapshot, checkpoint) in exactly the files the UNMODIFIED reference entry points read: train.ObjectNeRFSystem
(train.py:36-180), render_tools.editable_renderer.EditableRenderer (:53-332), utils.bbox_utils.BBoxRayHelper (:9-72).
Used by tests/test_dropin_reference.py to run those entry points once on the reference's own hot path and once over
object_nerf_b200.dropin."""
import json
import os
import sys
import types

import numpy as np
import torch
import yaml

from oracle import ref_loader as R

SCENE_ID = "scene0000_00"
N_MAX_VOXELS = 30000
LOSS = dict(color_loss_weight=1.0, depth_loss_weight=0.1, opacity_loss_weight=100.0, instance_color_loss_weight=1.0,
            instance_depth_loss_weight=0.1)


def purge_reference_modules():
    for name in list(sys.modules):
        if name == "train" or name.split(".")[0] in ("models", "render_tools", "utils", "datasets"):
            del sys.modules[name]


def write_scene(root):
    """-> (training config dict, paths)."""
    rng = np.random.default_rng(42)
    os.makedirs(os.path.join(root, "scans", SCENE_ID), exist_ok=True)
    os.makedirs(os.path.join(root, "bbox"), exist_ok=True)
    os.makedirs(os.path.join(root, "data"), exist_ok=True)
    os.makedirs(os.path.join(root, "ckpt"), exist_ok=True)
    # point cloud: a slab of random points, world units (scale_factor 2 -> +-0.55 in NeRF units)
    pts = rng.uniform([-1.1, -1.1, -0.5], [1.1, 1.1, 0.6], size=(6000, 3))
    pcd_path = os.path.join(root, "data", "pcd.ply")
    R.register_pointcloud(pcd_path, pts)
    # axis alignment: a small rotation about z plus an offset (4x4, row-major on one line as ScanNet writes it)
    ang = 0.2
    A = np.eye(4)
    A[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    A[:3, 3] = [0.1, -0.05, 0.02]
    with open(os.path.join(root, "scans", SCENE_ID, SCENE_ID + ".txt"), "w") as f:
        f.write("axisAlignment = " + " ".join(f"{v:.8f}" for v in A.reshape(-1)) + "\n")
    # instance boxes in the aligned frame: (cx, cy, cz, lx, ly, lz, instance id)
    boxes = np.array([[0.25, 0.10, 0.05, 0.7, 0.6, 0.5, 4], [-0.35, -0.20, 0.0, 0.5, 0.5, 0.6, 6]], dtype=np.float64)
    np.save(os.path.join(root, "bbox", SCENE_ID + "_bbox.npy"), boxes)
    # camera poses (Blender convention as the dataset stores them; load_frame_meta flips y / z back)
    frames = []
    for i, t in enumerate(np.linspace(0, 0.4, 3)):
        cam = np.array([-2.6 + t, 0.3, 0.35])
        fwd = -cam / np.linalg.norm(cam)
        right = np.cross(fwd, [0, 0, 1.0]); right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        Twc = np.eye(4)
        Twc[:3, :3] = np.stack([right, up, -fwd], 1)
        Twc[:3, 3] = cam
        stored = Twc.copy()
        stored[:3, :3] = stored[:3, :3] @ np.diag([1.0, -1.0, -1.0])      # undone by fix_rot (editable_renderer.py:93-96)
        frames.append({"idx": i, "transform_matrix": stored.tolist()})
    with open(os.path.join(root, "data", "transforms_full.json"), "w") as f:
        json.dump({"camera_angle_x": float(np.deg2rad(60.0)), "frames": frames}, f)
    conf = {
        "dataset_name": "scannet_base", "exp_name": "synthetic", "img_wh": [32, 24],
        "dataset_extra": {"near": 0.3, "far": 6.0, "scale_factor": 2.0, "scene_center": [0.0, 0.0, 0.0],
                          "root_dir": os.path.join(root, "data"), "bbox_dir": os.path.join(root, "bbox"),
                          "scans_dir": os.path.join(root, "scans"), "scene_id": SCENE_ID, "instance_id": [4, 6],
                          "pcd_path": pcd_path, "voxel_size": 0.1, "neighbor_marks": 3},
        "model": {"use_voxel_embedding": True, "N_freq_xyz": 10, "N_freq_dir": 4, "N_freq_voxel": 6, "D": 8, "W": 256,
                  "skips": [4], "N_scn_voxel_size": 16, "inst_D": 4, "inst_W": 128, "inst_skips": [2],
                  "N_obj_voxel_size": 8, "N_samples": 64, "N_importance": 64, "frustum_bound": 0.05, "use_disp": False,
                  "perturb": 0, "noise_std": 0, "N_max_objs": 64, "N_obj_code_length": 64, "N_max_voxels": N_MAX_VOXELS},
        "train": {"chunk": 32768, "batch_size": 256, "optimizer": "adam", "lr": 1e-3, "weight_decay": 0,
                  "progressive_train": False},
        "loss": dict(LOSS),
    }
    snap = os.path.join(root, "ckpt", "run_config_snapshot.yaml")
    with open(snap, "w") as f:
        yaml.safe_dump(conf, f)
    return conf, {"snapshot": snap, "ckpt": os.path.join(root, "ckpt", "last.ckpt")}


def fill_synthetic_weights(system, seed=7):
    """Reference-shaped random weights with sharpened density / colour heads, written through the state_dict (the
    keys are the checkpoint format both implementations share)."""
    from object_nerf_b200 import synthetic as S
    sd = system.state_dict()
    for prefix, s in (("nerf_coarse.", seed), ("nerf_fine.", seed + 1000)):
        w = S.make_weights(s, True, sigma_gain=8.0, sigma_bias=1.0, rgb_gain=16.0)
        for k, (W, b) in w.items():
            sd[prefix + S.REF_NAMES[k] + ".weight"] = W
            sd[prefix + S.REF_NAMES[k] + ".bias"] = b
    rng = np.random.default_rng(seed + 5)
    sd["code_library.embedding_instance.weight"] = torch.from_numpy(rng.standard_normal((64, 64)).astype(np.float32))
    sd["embedding_xyz.embedding_space_ftr.weight"] = torch.from_numpy(
        rng.standard_normal(tuple(sd["embedding_xyz.embedding_space_ftr.weight"].shape)).astype(np.float32))
    system.load_state_dict(sd, strict=True)
    return system


def training_batch(n=256, seed=3):
    from object_nerf_b200 import synthetic as S
    rng = np.random.default_rng(seed)
    rays = S.random_rays(seed + 1, n, near=0.15, far=3.0, cam_pos=(-1.3, 0.15, 0.17))
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.float32)))
    return {"rays": rays, "rgbs": f(rng.random((n, 3))), "depths": f(rng.uniform(0.3, 2.5, size=n)),
            "valid_mask": torch.from_numpy(rng.random(n) < 0.9), "instance_mask": torch.from_numpy(rng.random(n) < 0.5),
            "instance_mask_weight": f(np.where(rng.random(n) < 0.5, 1.0, 0.05)),
            "instance_ids": torch.from_numpy(rng.choice([4, 6], size=n)).view(n, 1),
            "pass_through_mask": torch.from_numpy(rng.random((n, 1)) < 0.5)}


def make_system(conf, device):
    """Import the reference's train.py (whatever `models.*` resolves to right now) and build its ObjectNeRFSystem."""
    import train
    system = train.ObjectNeRFSystem(R.to_attr(conf))
    system.train_dataset = types.SimpleNamespace(white_back=False, is_rays_in_bbox=lambda: False)   # generic_dataset.py API
    system.val_dataset = types.SimpleNamespace(white_back=False, is_rays_in_bbox=lambda: False)
    system = system.to(device)
    system.optimizer = torch.optim.Adam([p for p in system.parameters() if p.requires_grad], lr=1e-3)
    return train, system
