"""Import helper for the *real* reference (only usable where /root/reference exists, i.e. the
build container).  Used exclusively by tools/make_golden.py to pin oracle/ against the reference.
Never imported by the product, by bench.py or by the -m gpu tests.

Recipe follows SURVEY.md §8(c): stub the third-party modules the reference imports at module
scope but that the per-ray path never calls, and make `.cuda()` a no-op so the voxel embedding
constructor runs on CPU.
"""
import sys
import types

import numpy as np
import torch

REF_ROOT = "/root/reference"


class AttrDict(dict):
    """dict with attribute access (the reference indexes its config both ways)."""

    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


class _FakePcd:
    def __init__(self, pts):
        self.points = pts


_PCD_REGISTRY = {}


def register_pointcloud(path, pts):
    _PCD_REGISTRY[path] = np.asarray(pts, dtype=np.float64)


def install():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name in ("torch_optimizer", "matplotlib", "matplotlib.pyplot", "ipdb", "pytorch_lightning",
                 "imageio", "mcubes"):
        _stub(name)
    k = _stub("kornia")
    def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=None):
        """kornia is not installed: its published create_meshgrid restated (1 x H x W x 2, last dim = (x, y))."""
        xs = torch.linspace(0, width - 1, width)
        ys = torch.linspace(0, height - 1, height)
        if normalized_coordinates:
            xs = (xs / (width - 1) - 0.5) * 2
            ys = (ys / (height - 1) - 0.5) * 2
        base = torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1)
        return base.permute(1, 0, 2).unsqueeze(0)
    k.create_meshgrid = create_meshgrid  # datasets/ray_utils.py:2,17
    o3d = _stub("open3d")
    io = _stub("open3d.io")
    io.read_point_cloud = lambda p: _FakePcd(_PCD_REGISTRY[p])
    o3d.io = io
    # CPU-only container: the reference calls .cuda() unconditionally in the voxel helper
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


def default_model_config(use_voxel=True):
    # values of config/default_conf.yml:7-36
    return AttrDict(
        use_voxel_embedding=use_voxel, N_freq_xyz=10, N_freq_dir=4, N_freq_voxel=6, D=8, W=256,
        skips=[4], N_scn_voxel_size=16, inst_D=4, inst_W=128, inst_skips=[2], N_obj_voxel_size=8,
        N_samples=64, N_importance=64, frustum_bound=0.05, use_disp=False, perturb=1, noise_std=1,
        N_max_objs=64, N_obj_code_length=64, N_max_voxels=800000,
    )
