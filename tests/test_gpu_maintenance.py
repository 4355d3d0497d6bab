"""-m gpu: voxel grid maintenance with the device queries: raw trilinear features
(onerf_voxel_features) and the fused density query, against the reference-generated fixtures."""
import pytest
import torch

from oracle import onerf_oracle as O
from tests import cases, helpers
from tests.test_host_logic_cpu import _maint_embedding, _oracle_grid

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def test_voxel_features_match_oracle():
    from object_nerf_b200 import engine
    emb, _ = _maint_embedding()
    grid = _oracle_grid(emb)
    g = torch.Generator().manual_seed(3)
    lo = -emb.voxel_offset
    hi = lo + emb.voxel_size * emb.voxel_shape.float()
    xyz = lo - 0.3 + (hi - lo + 0.6) * torch.rand(5000, 3, generator=g)      # inside, in empty cells and outside
    want = O.voxel_features(xyz, grid)
    emb = emb.to(DEV)
    got = engine.voxel_features(xyz.to(DEV), emb.grid_buffers()).cpu()
    assert tuple(got.shape) == (5000, 24)
    assert torch.allclose(got, want, rtol=0, atol=1e-6), (got - want).abs().max().item()
    assert ((want == 0).all(1) == (got == 0).all(1)).all() and (want != 0).any()


def test_voxel_subdivision_on_device_matches_reference_golden(golden):
    gold = golden("maint_subdivision")
    emb, _ = _maint_embedding()
    emb = emb.to(DEV)
    n_after = emb.voxel_subdivision()
    assert n_after == int(gold["subdiv|table_rows"].shape[0])
    assert torch.equal(emb.voxel_size.cpu(), gold["subdiv|voxel_size"])
    assert torch.equal(emb.voxel_shape.cpu(), gold["subdiv|voxel_shape"])
    assert torch.equal(emb.voxel_occupancy.cpu(), gold["subdiv|voxel_occupancy"].bool())
    assert torch.equal(emb.voxel_idx_map.cpu(), gold["subdiv|voxel_idx_map"])
    rows = emb.embedding_space_ftr.weight.detach()[:n_after].cpu()
    assert torch.allclose(rows, gold["subdiv|table_rows"], rtol=0, atol=1e-6)


def test_self_pruning_on_device_matches_reference_golden(golden):
    gold = golden("maint_pruning")
    emb, inp = _maint_embedding()
    n_occu = int(gold["n_before"])
    model = helpers.make_model(inp["weights"], True, DEV)
    emb = emb.to(DEV)
    rand = [r.to(DEV) for r in cases.maint_rand((n_occu + 31) // 32)]
    # fp32 arithmetic: the threshold sits in a 4 % gap of the per-voxel maximum alphas (tests/cases.py)
    n_pruned = emb.self_pruning_empty_voxels(model, max_alpha_th=cases.MAINT_CASE["max_alpha_th"], precision="fp32", _rand=rand)
    assert n_pruned == n_occu - int(gold["pruned|voxel_occupancy"].sum())
    assert torch.equal(emb.voxel_occupancy.cpu(), gold["pruned|voxel_occupancy"].bool())
    assert torch.equal(emb.voxel_idx_map.cpu(), gold["pruned|voxel_idx_map"])
