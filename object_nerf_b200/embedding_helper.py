"""Embedding modules with the reference's names / buffers (reference models/embedding_helper.py).

`Embedding` is a marker for plain positional encoding; `EmbeddingVoxel` owns the sparse voxel grid
(feature table + dense index map + metadata buffers, same state_dict keys as the reference) and builds
it from a point cloud at construction time (cold path).  The per-sample encoding itself runs inside the
fused CUDA kernels; calling these modules directly uses the stand-alone encode kernel.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import engine


class Embedding(nn.Module):
    """[x, sin(2^k x), cos(2^k x)]_k, reference :40-74.  Supported: in_channels=3 with 10 or 4 octaves."""

    def __init__(self, in_channels, N_freqs, logscale=True):
        super().__init__()
        if not logscale:
            raise RuntimeError("only log-scale frequency bands are built")
        self.N_freqs, self.in_channels = N_freqs, in_channels
        self.out_channels = in_channels * (2 * N_freqs + 1)

    def forward(self, x):
        if self.in_channels == 3 and self.N_freqs == 10 and x.is_cuda:
            return engine.encode(x.reshape(-1, 3), None)[0].reshape(*x.shape[:-1], 63)
        raise NotImplementedError("stand-alone Embedding.forward is only built for PE10 of CUDA xyz; "
                                  "direction encoding happens inside the fused field kernel")


class EmbeddingVoxel(nn.Module):
    def __init__(self, channels, N_freqs, max_voxels, dataset_extra_config, points=None):
        super().__init__()
        if (channels, N_freqs) != (24, 6):
            raise RuntimeError("object_nerf_b200 kernels are built for 24 voxel channels with PE 6")
        self.channels = channels
        self.instance_ftr_C = 8
        self.embedding_space_ftr = nn.Embedding(max_voxels, channels)
        self.conf = dataset_extra_config
        self.set_pointclouds(dataset_extra_config, points)

    # ---- cold path: grid construction (reference :86-200) ----
    def set_pointclouds(self, conf, points=None):
        if points is None:
            import open3d as o3d  # only needed to read the .ply the dataset config names
            points = np.asarray(o3d.io.read_point_cloud(conf["pcd_path"]).points)
        scale = conf["scale_factor"]
        pts = torch.from_numpy((np.asarray(points) - np.array(conf["scene_center"])) / scale).float()
        vsize = torch.scalar_tensor(conf["voxel_size"] / scale)
        lo, hi = pts.min(0)[0], pts.max(0)[0]
        self.register_buffer("voxel_size", vsize)
        self.register_buffer("bounds", torch.stack([lo, hi]))
        self.register_buffer("voxel_offset", -lo)
        shape = [int(((hi[i] - lo[i]) / vsize).int().item()) + 3 for i in range(3)]
        self.register_buffer("voxel_shape", torch.tensor(shape))
        self.register_buffer("voxel_count", torch.scalar_tensor(int(np.prod(shape))))
        occ = torch.zeros(shape, dtype=torch.bool)
        q = ((pts + self.voxel_offset) / vsize).round().long()
        ok = ((q >= 0) & (q < torch.tensor(shape))).all(1)
        q = q[ok]
        occ[q[:, 0], q[:, 1], q[:, 2]] = True
        # mark the neighbourhood of every occupied cell (box dilation)
        k = int(conf["neighbor_marks"])
        occ = F.max_pool3d(occ[None, None].float(), kernel_size=k, stride=1, padding=(k - 1) // 2)[0, 0] > 0
        self.register_buffer("voxel_occupancy", occ)
        self.generate_voxel_idx_map()

    def generate_voxel_idx_map(self):
        occ = self.voxel_occupancy
        cells = torch.nonzero(occ)
        if cells.shape[0] > self.embedding_space_ftr.num_embeddings:
            raise RuntimeError("more occupied voxels than N_max_voxels")
        idx = torch.full(tuple(occ.shape), -1, dtype=torch.long, device=occ.device)
        idx[cells[:, 0], cells[:, 1], cells[:, 2]] = torch.arange(cells.shape[0], device=occ.device)
        self.register_buffer("voxel_idx_map", idx)

    def grid_buffers(self) -> engine.GridBuffers:
        return engine.GridBuffers.from_module(self)

    def forward(self, xyz):
        scene, obj = engine.encode(xyz.reshape(-1, 3), self.grid_buffers())
        return scene, obj
