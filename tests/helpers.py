"""Test helpers: build the package's parameter containers from the synthetic scenes of tests/synth.py."""
import torch
from torch import nn

from tests import synth

REF_NAMES = {  # oracle layout -> reference attribute names (models/nerf_model.py:41-58,77-95)
    **{f"scene.l{i}": f"xyz_encoding_{i+1}.0" for i in range(8)},
    "scene.final": "xyz_encoding_final", "scene.sigma": "sigma", "scene.dir": "dir_encoding.0",
    "scene.rgb": "rgb.0",
    **{f"obj.l{i}": f"instance_encoding_{i+1}.0" for i in range(4)},
    "obj.final": "instance_encoding_final.0", "obj.sigma": "instance_sigma",
    "obj.dir": "inst_dir_encoding.0", "obj.rgb": "inst_rgb.0",
}


class Cfg(dict):
    __getattr__ = dict.__getitem__


def model_config(use_voxel=True):
    return Cfg(use_voxel_embedding=use_voxel, N_freq_xyz=10, N_freq_dir=4, N_freq_voxel=6, D=8, W=256,
               skips=[4], N_scn_voxel_size=16, inst_D=4, inst_W=128, inst_skips=[2], N_obj_voxel_size=8,
               N_max_objs=64, N_obj_code_length=64, N_max_voxels=800000)


def make_model(w, use_voxel, device):
    from object_nerf_b200 import ObjectNeRF
    m = ObjectNeRF(model_config(use_voxel))
    sd = {}
    for k, (W, b) in w.items():
        sd[REF_NAMES[k] + ".weight"] = W
        sd[REF_NAMES[k] + ".bias"] = b
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


class GridModule(nn.Module):
    """Stands in for EmbeddingVoxel with an injected grid: the same buffers / parameter the kernels read."""

    def __init__(self, g):
        super().__init__()
        self.embedding_space_ftr = nn.Embedding.from_pretrained(g["table"].clone(), freeze=False)
        self.register_buffer("voxel_idx_map", g["idx_map"].clone())
        self.register_buffer("voxel_offset", g["offset"].clone())
        self.register_buffer("voxel_size", g["voxel_size"].clone())
        self.register_buffer("voxel_shape", g["shape"].clone())


class CodeLib(nn.Module):
    def __init__(self, table):
        super().__init__()
        self.embedding_instance = nn.Embedding.from_pretrained(table.clone(), freeze=False)


def psnr(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float("inf") if mse == 0 else -10.0 * torch.log10(torch.tensor(mse)).item()
