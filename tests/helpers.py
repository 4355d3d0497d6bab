"""Test helpers: build the package's parameter containers from the synthetic scenes of tests/synth.py."""
import torch
from torch import nn

from tests import synth

from object_nerf_b200.synthetic import REF_NAMES, Cfg, GridModule, make_model, model_config  # noqa: E402,F401


class CodeLib(nn.Module):
    def __init__(self, table):
        super().__init__()
        self.embedding_instance = nn.Embedding.from_pretrained(table.clone(), freeze=False)


def psnr(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float("inf") if mse == 0 else -10.0 * torch.log10(torch.tensor(mse)).item()


# ------------------------------------------------------------------------------------------------
# training workspace ("atoms") helpers: Python mirror of object_nerf_b200/csrc/layout.h (TrainLayout)
# ------------------------------------------------------------------------------------------------
ATOM_BYTES = 16384
MASK_WORDS = 88
ACT_ATOMS_VOXEL = [6, 4, 4, 4, 4, 4, 4, 4, 4, 4, 2, 2, 2, 2, 2, 2, 1]


def train_layout(use_voxel, n_samples):
    n_tiles = (n_samples + 127) // 128
    aw = list(ACT_ATOMS_VOXEL)
    if not use_voxel:
        aw[0] = 1
    off, act_off, dz_off = 0, [], []
    for a in aw:
        act_off.append(off)
        off += a * n_tiles * ATOM_BYTES
    for a in aw[1:]:
        dz_off.append(off)
        off += a * n_tiles * ATOM_BYTES
    off += ATOM_BYTES
    mask_off = off
    off += n_tiles * MASK_WORDS * 128 * 4
    return dict(n_tiles=n_tiles, act_atoms=aw, dz_atoms=aw[1:], act_off=act_off, dz_off=dz_off, mask_off=mask_off,
                total=(off + 1023) // 1024 * 1024)


def _swizzle_index(device):
    r = torch.arange(128, device=device).view(128, 1)
    c = torch.arange(8, device=device).view(1, 8)
    return (c ^ (r & 7))          # (128, 8): physical chunk <-> logical chunk (involution)


def to_atoms(mat):
    """(B, 64 * A) float matrix (B multiple of 128) -> bf16 atom images, flat int16-viewable tensor in slot order
    [tile][atom][row 128][chunk 8 (swizzled)][8]."""
    B, Wd = mat.shape
    T, A = B // 128, Wd // 64
    x = mat.to(torch.bfloat16).view(T, 128, A, 8, 8).permute(0, 2, 1, 3, 4).contiguous()      # T, A, row, chunk, 8
    idx = _swizzle_index(mat.device).view(1, 1, 128, 8, 1).expand(T, A, 128, 8, 8)
    return torch.gather(x, 3, idx).contiguous()


def from_atoms(buf_u8, off, n_tiles, atoms):
    """Read a slot back: uint8 workspace tensor -> (n_tiles * 128, 64 * atoms) float32."""
    n = n_tiles * atoms * ATOM_BYTES
    x = buf_u8[off:off + n].view(torch.bfloat16).view(n_tiles, atoms, 128, 8, 8)
    idx = _swizzle_index(buf_u8.device).view(1, 1, 128, 8, 1).expand(n_tiles, atoms, 128, 8, 8)
    x = torch.gather(x, 3, idx)
    return x.permute(0, 2, 1, 3, 4).reshape(n_tiles * 128, atoms * 64).float()


def write_atoms(buf_u8, off, mat):
    a = to_atoms(mat)
    n = a.numel() * 2
    buf_u8[off:off + n] = a.view(torch.uint8).reshape(-1)


def read_masks(buf_u8, layout):
    """-> (n_tiles, 88, 128) int64 mask words."""
    n = layout["n_tiles"] * MASK_WORDS * 128 * 4
    return buf_u8[layout["mask_off"]:layout["mask_off"] + n].view(torch.int32).view(layout["n_tiles"], MASK_WORDS, 128).to(torch.int64) & 0xffffffff


def aligned_u8(nbytes, device, fill=None):
    t = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-t.data_ptr()) % 1024
    t = t[off:off + nbytes]
    if fill is not None:
        t.fill_(fill)
    return t
