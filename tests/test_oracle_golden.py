"""Pin oracle/ against outputs of the real reference (fixtures made by tools/make_golden.py).
CPU only.  The oracle uses the reference's own array library and op order, so the comparison is
bit-exact (torch.equal) wherever the reference's kernels are deterministic."""
import pytest
import torch

from oracle import onerf_oracle as O
from tests import cases, synth


def grid_obj(g):
    return O.VoxelGrid(g["offset"], g["voxel_size"], g["shape"].tolist(), g["idx_map"], g["table"])


def assert_same(a, b, name, exact=True, tol=0.0):
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if exact:
        assert torch.equal(a, b), (name, (a - b).abs().max().item())
    else:
        assert torch.allclose(a, b, rtol=tol, atol=tol), (name, (a - b).abs().max().item())


def test_posenc(golden):
    g = golden("stage_posenc")
    x = cases.stage_inputs()["posenc_x"]
    assert_same(O.posenc(x, 10), g["pe10"], "pe10")
    assert_same(O.posenc(x, 4), g["pe4"], "pe4")


def test_voxel_embed(golden):
    g = golden("stage_voxel")
    grid = grid_obj(synth.make_grid(**cases.GRID_KW))
    s, o = O.voxel_embed(cases.stage_inputs()["voxel_xyz"], grid)
    # trilinear blend: reference sums a stacked (8,N,C) tensor; association may differ in the last ulp,
    # and PE(2^5 * f) amplifies an ulp of f by 32.
    assert_same(s, g["scene_in"], "scene_in", exact=True)
    assert_same(o, g["obj_in"], "obj_in", exact=True)
    assert (s[:, 208:] == g["scene_in"][:, 208:]).all()  # classic PE part is exact


@pytest.mark.parametrize("use_voxel", [True, False])
def test_mlp(golden, use_voxel):
    g = golden(f"stage_mlp_{'voxel' if use_voxel else 'plain'}")
    si = cases.stage_inputs()
    w = synth.make_weights(11, use_voxel, sigma_gain=8.0, sigma_bias=1.0)
    ex = si["emb_xyz_v"] if use_voxel else si["emb_xyz_p"]
    sigma, rgb = O.scene_mlp(w, ex, si["emb_dir"])
    isg, irgb = O.object_mlp(w, ex, si["obj_voxel"] if use_voxel else None, si["obj_code"], si["emb_dir"])
    for a, k in ((sigma, "sigma"), (rgb, "rgb"), (isg, "inst_sigma"), (irgb, "inst_rgb")):
        assert_same(a, g[k], k, exact=True)


def test_sample_pdf(golden):
    g = golden("stage_sample_pdf")
    si = cases.stage_inputs()
    assert_same(O.sample_pdf(si["pdf_bins"], si["pdf_weights"], 64, det=True), g["det"], "det")
    assert_same(O.sample_pdf(si["pdf_bins"], si["pdf_weights"], 64, det=False, u=si["pdf_u"]), g["rnd"], "rnd")


@pytest.mark.parametrize("name", list(cases.RENDER_CASES))
def test_render_rays(golden, name):
    c = cases.RENDER_CASES[name]
    g = golden("render_" + name)
    inp = cases.build_render_case(c)
    grid = grid_obj(inp["grid"]) if inp["grid"] is not None else None
    out = O.render_rays(inp["weights"], grid, inp["rays"], inp["codes"], n_samples=c["n_samples"],
                        use_disp=c["use_disp"], perturb=c["perturb"], noise_std=c["noise_std"],
                        n_importance=c["n_importance"], white_back=c["white_back"],
                        forward_instance=c["forward_instance"], frustum_bound_th=c["frustum_bound_th"],
                        pass_through_mask=inp["pass_through_mask"], rays_in_bbox=c["rays_in_bbox"],
                        is_eval=c["is_eval"], rand=inp["rand"])
    assert set(out) == set(g), (sorted(out), sorted(g))
    for k in g:
        assert_same(out[k], g[k], k, exact=True)


@pytest.mark.parametrize("name", list(cases.MULTI_CASES))
def test_render_rays_multi(golden, name):
    c = cases.MULTI_CASES[name]
    g = golden("multi_" + name)
    inp = cases.build_multi_case(c)
    boxes = [cases.box_affine(b) for b in inp["boxes"]]
    out = O.render_rays_multi(inp["weights"], grid_obj(inp["grid"]), inp["code_table"], inp["rays_list"],
                              c["obj_ids"], n_samples=c["n_samples"], n_importance=c["n_importance"],
                              white_back=c["white_back"], skip_boxes=boxes)
    assert set(out) == set(g), (sorted(out), sorted(g))
    for k in g:
        assert_same(out[k], g[k], k, exact=True)


def test_training_step_gradients_match_reference(golden):
    """Oracle autograd (torch CPU) vs the reference's backward through render_rays + TotalLoss
    (fixture: loss, and per parameter tensor the L2 norm, the sum and 256 sampled entries)."""
    g = golden("grad_train_step")
    c = cases.GRAD_CASE
    inp = cases.build_grad_case()
    leaves = {}

    def leaf(name, t):
        t = t.clone().requires_grad_(True)
        leaves[name] = t
        return t

    weights = {typ: {k: (leaf(f"{typ}.{helpers_name(k)}.weight", W), leaf(f"{typ}.{helpers_name(k)}.bias", b))
                     for k, (W, b) in w.items()} for typ, w in inp["weights"].items()}
    table = leaf("voxel", inp["grid"]["table"])
    code_table = leaf("codes", inp["code_table"])
    gd = inp["grid"]
    grid = O.VoxelGrid(gd["offset"], gd["voxel_size"], gd["shape"].tolist(), gd["idx_map"], table)
    codes = code_table[inp["instance_ids"].view(-1)]
    out = O.render_rays(weights, grid, inp["rays"], codes, n_samples=c["n_samples"], perturb=c["perturb"],
                        noise_std=c["noise_std"], n_importance=c["n_importance"], frustum_bound_th=c["frustum_bound_th"],
                        pass_through_mask=inp["pass_through_mask"], is_eval=False, rand=inp["rand"])
    loss = cases.total_loss(out, inp["batch"])
    assert abs(loss.item() - g["loss"].item()) <= 1e-5 * abs(g["loss"].item())
    loss.backward()
    for name, t in leaves.items():
        gr = t.grad.reshape(-1)
        ref_norm = g[name + "|norm"].item()
        assert abs(gr.norm().item() - ref_norm) <= 1e-4 * max(ref_norm, 1e-6), name
        idx = cases.sample_indices(name, gr.numel())
        assert torch.allclose(gr[idx], g[name + "|samples"], rtol=1e-3, atol=1e-5 * max(ref_norm, 1e-6)), name
    nz = torch.nonzero(table.grad.abs().sum(1)).view(-1)
    assert torch.equal(nz, g["voxel|nonzero_rows"])


def helpers_name(k):
    from tests.helpers import REF_NAMES
    return REF_NAMES[k]


# ------------------------------------------------------------------------------------------------
# camera rays / per-object ray assembly
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.CAMERA_CASES))
def test_camera_rays_match_reference_golden(golden, name):
    inp = cases.build_camera_case(cases.CAMERA_CASES[name])
    gold = golden("rays_" + name)
    directions = O.ray_directions(inp["H"], inp["W"], inp["focal"])
    assert torch.equal(directions, gold["directions"])
    rays_o, rays_d = O.get_rays(directions, inp["c2w"])
    assert torch.equal(rays_o, gold["rays_o"]) and torch.equal(rays_d, gold["rays_d"])


@pytest.mark.parametrize("name", list(cases.BBOX_CASES))
def test_ray_bbox_intersections_match_reference_golden(golden, name):
    inp = cases.build_bbox_case(cases.BBOX_CASES[name])
    gold = golden("rays_" + name)
    mask, near, far = O.ray_bbox_intersections(inp["rays_o"], inp["rays_d"], inp["pose_avg"], inp["axis_align_mat"],
                                               inp["bbox_bounds"], inp["scale_factor"], inp["bbox_enlarge"])
    assert torch.equal(mask, gold["mask"].bool())
    assert torch.equal(near, gold["near"]) and torch.equal(far, gold["far"])
    # the cases exercise every rule of the slab test
    d = inp["rays_d"]
    assert (d == 0).any() and mask.any() and (~mask).any()


# ------------------------------------------------------------------------------------------------
# training loss
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.LOSS_CASES))
def test_total_loss_matches_reference_golden(golden, name):
    maps, batch = cases.build_loss_case(cases.LOSS_CASES[name])
    gold = golden(name)
    maps = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
    loss_sum, terms = O.total_loss(maps, batch, cases.LOSS_CONF)
    assert torch.equal(loss_sum.detach(), gold["loss_sum"])
    assert {"term|" + k for k in terms} == {k for k in gold if k.startswith("term|")}
    for k, v in terms.items():
        assert torch.equal(v.detach(), gold["term|" + k]), k
    loss_sum.backward()
    for k, v in maps.items():
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        assert torch.equal(g, gold["grad|" + k]), k
