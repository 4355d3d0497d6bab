"""Parameter container with the reference's ObjectNeRF attribute names / state_dict keys
(reference models/nerf_model.py:18-95), so checkpoints and optimizers are interchangeable.

It holds weights only.  The arithmetic lives in the CUDA library: render_rays() packs these
parameters (engine.packed_for) and runs the fused kernels.
"""
from __future__ import annotations

from torch import nn


def _cfg(cfg, key, default=None):
    try:
        return cfg[key]
    except (KeyError, TypeError):
        return getattr(cfg, key, default)


def _act_linear(k, n, act):
    return nn.Sequential(nn.Linear(k, n), act)


class ObjectNeRF(nn.Module):
    def __init__(self, model_config):
        super().__init__()
        self.model_config = model_config
        self.use_voxel_embedding = bool(_cfg(model_config, "use_voxel_embedding", True))
        D, W = _cfg(model_config, "D"), _cfg(model_config, "W")
        skips = list(_cfg(model_config, "skips"))
        inst_D, inst_W = _cfg(model_config, "inst_D"), _cfg(model_config, "inst_W")
        inst_skips = list(_cfg(model_config, "inst_skips"))
        nfx, nfd = _cfg(model_config, "N_freq_xyz"), _cfg(model_config, "N_freq_dir")
        if (D, W, skips, inst_D, inst_W, inst_skips, nfx, nfd) != (8, 256, [4], 4, 128, [2], 10, 4):
            raise RuntimeError("object_nerf_b200 kernels are built for the default architecture only "
                               "(D=8, W=256, skips=[4], inst_D=4, inst_W=128, inst_skips=[2], PE 10/4)")
        self.D, self.W, self.skips = D, W, skips
        self.inst_D, self.inst_W, self.inst_skips = inst_D, inst_W, inst_skips
        self.N_freq_xyz, self.N_freq_dir = nfx, nfd
        vox_scene = vox_obj = 0
        if self.use_voxel_embedding:
            self.N_freq_voxel = _cfg(model_config, "N_freq_voxel")
            self.N_scn_voxel_size = _cfg(model_config, "N_scn_voxel_size", 0)
            n_obj_vox = _cfg(model_config, "N_obj_voxel_size", 0)
            if (self.N_freq_voxel, self.N_scn_voxel_size, n_obj_vox) != (6, 16, 8):
                raise RuntimeError("object_nerf_b200 kernels are built for 16+8 voxel channels with PE 6")
            vox_scene = self.N_scn_voxel_size * (1 + 2 * self.N_freq_voxel)
            vox_obj = n_obj_vox * (1 + 2 * self.N_freq_voxel)
        n_code = _cfg(model_config, "N_obj_code_length")
        if n_code != 64:
            raise RuntimeError("object_nerf_b200 kernels are built for 64-long object codes")
        self.in_channels_xyz = 3 * (1 + 2 * nfx) + vox_scene
        self.in_channels_dir = 3 * (1 + 2 * nfd)
        self.inst_channel_in = self.in_channels_xyz + n_code + vox_obj
        self.activation = nn.LeakyReLU(inplace=True)
        act = self.activation
        # creation order below fixes the state_dict key order and the RNG consumption of default init
        for i in range(D):
            k = self.in_channels_xyz if i == 0 else (W + self.in_channels_xyz if i in skips else W)
            setattr(self, f"xyz_encoding_{i+1}", _act_linear(k, W, act))
        self.xyz_encoding_final = nn.Linear(W, W)
        self.sigma = nn.Linear(W, 1)
        self.rgb = nn.Sequential(nn.Linear(W // 2, 3), nn.Sigmoid())
        self.dir_encoding = _act_linear(W + self.in_channels_dir, W // 2, act)
        for i in range(inst_D):
            k = self.inst_channel_in if i == 0 else (inst_W + self.inst_channel_in if i in inst_skips else inst_W)
            setattr(self, f"instance_encoding_{i+1}", _act_linear(k, inst_W, act))
        self.instance_encoding_final = nn.Sequential(nn.Linear(inst_W, inst_W))
        self.instance_sigma = nn.Linear(inst_W, 1)
        self.inst_dir_encoding = _act_linear(inst_W + self.in_channels_dir, inst_W // 2, act)
        self.inst_rgb = nn.Sequential(nn.Linear(inst_W // 2, 3), nn.Sigmoid())

    def _density(self, inputs, obj_code):
        src = getattr(inputs.get("emb_xyz"), "_onerf_src", None) if isinstance(inputs, dict) else None
        if src is None:
            raise NotImplementedError(
                "ObjectNeRF is a weight container: the MLP runs fused with the encoding inside render_rays() / "
                "render_rays_multi() / query_sigma().  forward(..., sigma_only=True) is supported on inputs produced by "
                "EmbeddingVoxel.forward(xyz) (they carry their positions); arbitrary pre-embedded features are not.")
        from . import rendering
        pts, emb = src
        return rendering.query_sigma(self, emb, pts, obj_code=obj_code)[:, None]

    def forward(self, inputs, sigma_only=False):
        """Reference :97-121.  Only the density query (`sigma_only=True`: tools/extract_mesh.py:104-107, the voxel pruning
        of embedding_helper.py:219-225) is served here, through the fused kernel; colours come from render_rays()."""
        if not sigma_only:
            raise NotImplementedError("per-sample colours are produced inside render_rays() / render_rays_multi()")
        return {"sigma": self._density(inputs, None)}

    def forward_instance(self, inputs, sigma_only=False):
        """Reference :123-152, density only (tools/extract_mesh.py:95-103): inputs["obj_code"] holds one code per point,
        all rows equal (the reference looks the same id up for every point)."""
        if not sigma_only:
            raise NotImplementedError("per-sample colours are produced inside render_rays() / render_rays_multi()")
        code = inputs["obj_code"]
        return {"inst_sigma": self._density(inputs, code[0] if code.dim() == 2 else code)}
