"""Full-size checks (BASELINE.json configs[1]: 640x480 frame, 64 + 64 samples, two-branch voxel model) through
size-independent properties, plus the north_star's acceptance criterion: PSNR delta < 0.05 dB between the tensor-core
(bf16) render and the fp32 render against the same ground truth.  The fp32 kernel is itself pinned to the reference
goldens at small sizes (test_gpu_parity.py); the CPU oracle cannot finish a full frame in seconds."""
import numpy as np
import pytest
import torch

import bench
from tests import helpers

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def scene():
    from object_nerf_b200 import Embedding
    sc = bench.build_scene(DEV)
    models = {k: helpers.make_model(w, True, DEV) for k, w in sc["weights"].items()}
    emb = helpers.GridModule(sc["grid"]).to(DEV)
    return dict(models=models, embeddings={"xyz": emb, "dir": Embedding(3, 4)}, rays=sc["rays"].to(DEV),
                codes=sc["codes"].to(DEV))


def render(scene, sl, precision, **kw):
    from object_nerf_b200 import render_rays
    args = dict(N_samples=64, perturb=0, noise_std=0, N_importance=64, is_eval=True, precision=precision)
    args.update(kw)
    with torch.no_grad():
        return render_rays(scene["models"], scene["embeddings"], scene["rays"][sl], embedding_instance=scene["codes"][sl], **args)


def test_full_frame_invariants(scene):
    n = bench.N_RAYS
    outs = [render(scene, slice(i, i + 65536), "bf16") for i in range(0, n, 65536)]
    out = {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}
    assert out["rgb_fine"].shape == (n, 3) and out["weights_fine"].shape == (n, 128)
    for typ in ("coarse", "fine"):
        z, w = out[f"z_vals_{typ}"], out[f"weights_{typ}"]
        assert torch.isfinite(w).all() and torch.isfinite(out[f"rgb_{typ}"]).all()
        assert (z[:, 1:] >= z[:, :-1]).all(), "depths must be sorted"
        assert (w >= 0).all()
        assert torch.allclose(w.sum(1), out[f"opacity_{typ}"], atol=2e-5)          # opacity = sum of weights
        assert (out[f"opacity_{typ}"] <= 1 + 1e-4).all() and (out[f"opacity_instance_{typ}"] <= 1 + 1e-4).all()
        assert (out[f"rgb_{typ}"] >= -1e-5).all() and (out[f"rgb_{typ}"] <= 1 + 1e-4).all()
        near, far = scene["rays"][:, 6], scene["rays"][:, 7]
        assert (z[:, 0] >= near - 1e-5).all() and (z[:, -1] <= far + 1e-5).all()
        d = out[f"depth_{typ}"]
        assert (d >= -1e-5).all() and (d <= far * (1 + 1e-4) + 1e-4).all()
        # object branch is composited on white: rgb_instance -> 1 where the object is absent
        empty = out[f"opacity_instance_{typ}"] < 1e-4
        if empty.any():
            assert torch.allclose(out[f"rgb_instance_{typ}"][empty], torch.ones_like(out[f"rgb_instance_{typ}"][empty]), atol=2e-4)
    # the fine depths contain the coarse depths (models/rendering.py:313)
    zc, zf = out["z_vals_coarse"][:4096], out["z_vals_fine"][:4096]
    pos = torch.searchsorted(zf.contiguous(), zc.contiguous())
    assert torch.equal(torch.gather(zf, 1, pos.clamp(max=127)), zc)


def test_chunk_invariance_and_determinism(scene):
    """Rays are independent: rendering a block in one call or in ragged pieces gives bit-identical maps; two runs of
    the same call are bit-identical (no atomics / races on the forward path)."""
    sl = slice(1000, 1000 + 20000)
    a = render(scene, sl, "bf16")
    b = render(scene, sl, "bf16")
    parts = [render(scene, slice(1000 + i, 1000 + j), "bf16") for i, j in ((0, 7), (7, 8200), (8200, 20000))]
    for k in ("rgb_fine", "depth_fine", "weights_fine", "rgb_instance_fine", "opacity_instance_coarse", "z_vals_fine"):
        assert torch.equal(a[k], b[k]), k
        assert torch.equal(a[k], torch.cat([p[k] for p in parts], 0)), k


def test_bf16_render_within_psnr_tolerance_of_fp32(scene):
    """north_star: 'output within stated tolerance of the reference (PSNR delta < 0.05 dB)'.  Ground truth = the fp32
    render plus 30 dB gaussian noise (a trained model's typical PSNR); the tensor-core render must score within
    0.05 dB of the fp32 render against it, on a 32 768-ray subsample of the frame."""
    idx = torch.arange(0, bench.N_RAYS, bench.N_RAYS // 32768, device=DEV)[:32768]
    sub = dict(scene, rays=scene["rays"][idx].contiguous(), codes=scene["codes"][idx].contiguous())
    hi = render(sub, slice(None), "fp32")
    lo = render(sub, slice(None), "bf16")
    g = torch.Generator(device=DEV).manual_seed(0)
    for k in ("rgb_fine", "rgb_instance_fine", "rgb_coarse"):
        gt = hi[k] + torch.randn(hi[k].shape, device=DEV, generator=g) * (10 ** (-30 / 20))
        p_hi, p_lo = helpers.psnr(hi[k], gt), helpers.psnr(lo[k], gt)
        assert abs(p_hi - p_lo) < 0.05, (k, p_hi, p_lo)
        assert helpers.psnr(lo[k], hi[k]) > 45.0, (k, helpers.psnr(lo[k], hi[k]))
    assert (lo["depth_fine"] - hi["depth_fine"]).abs().mean().item() < 5e-3


def test_training_mode_device_rng_statistics(scene):
    """perturb / noise_std with the library's Philox RNG (no injected buffers): jittered depths stay stratified and
    different calls draw different samples; the render stays finite."""
    sl = slice(0, 4096)
    a = render(scene, sl, "bf16", perturb=1.0, noise_std=1.0, is_eval=False, frustum_bound_th=0.025)
    b = render(scene, sl, "bf16", perturb=1.0, noise_std=1.0, is_eval=False, frustum_bound_th=0.025)
    assert not torch.equal(a["z_vals_coarse"], b["z_vals_coarse"])
    z = a["z_vals_coarse"]
    assert (z[:, 1:] >= z[:, :-1]).all()
    assert torch.isfinite(a["rgb_fine"]).all() and torch.isfinite(a["rgb_instance_fine"]).all()
    # each jittered depth stays inside its stratum [mid_{i-1}, mid_i] of the regular grid (models/rendering.py:268-277)
    near, far = scene["rays"][sl, 6:7], scene["rays"][sl, 7:8]
    t = torch.linspace(0, 1, 64, device=DEV)
    base = near * (1 - t) + far * t
    mid = 0.5 * (base[:, 1:] + base[:, :-1])
    assert (z[:, 1:-1] >= mid[:, :-1] - 1e-5).all() and (z[:, 1:-1] <= mid[:, 1:] + 1e-5).all()


# ------------------------------------------------------------------------------------------------
# parity at the benchmark configuration, directly against the reference (oracle/_ref: the unmodified reference
# byte-compiled at build time) or, where that is absent, the oracle port pinned to it
# ------------------------------------------------------------------------------------------------
def test_bench_scene_render_matches_reference_directly(scene):
    """2 048 rays of the bench frame (the bench's own scene, weights and sample counts), bf16 and fp32, against the
    reference's render_rays on the host.  Tolerances are the measured errors plus margin (profiles/r02_parity.md)."""
    sc = bench.build_scene()
    kind, ref_render = bench.reference_renderer(sc)
    sel = torch.linspace(0, bench.N_RAYS - 1, 2048).long()
    want = ref_render(sc["rays"][sel], sc["codes"][sel])
    sub = dict(scene, rays=scene["rays"][sel.to(DEV)].contiguous(), codes=scene["codes"][sel.to(DEV)].contiguous())
    assert want["rgb_fine"].std().item() > 0.05, "the bench scene must have structure for this test to mean anything"
    tol = {"fp32": dict(rgb=2e-5, depth=2e-5, inst=5e-5, psnr=100.0), "bf16": dict(rgb=3e-3, depth=5e-3, inst=1e-2, psnr=65.0)}
    for precision in ("fp32", "bf16"):
        got = {k: v.cpu() for k, v in render(sub, slice(None), precision).items()}
        t = tol[precision]
        assert helpers.psnr(got["rgb_fine"], want["rgb_fine"]) > t["psnr"], (precision, helpers.psnr(got["rgb_fine"], want["rgb_fine"]))
        for k, lim in (("rgb_fine", t["rgb"]), ("rgb_coarse", t["rgb"]), ("depth_fine", t["depth"]),
                       ("rgb_instance_fine", t["inst"]), ("opacity_instance_fine", t["inst"]), ("depth_instance_fine", t["inst"])):
            err = (got[k] - want[k]).abs()
            # importance samples at u = 1 are a knife edge of the reference itself (test_gpu_parity.py): bound their share
            assert (err > lim).float().mean().item() < 5e-3, (precision, k, err.max().item())
        # PSNR delta against a 30 dB "photograph" of the reference render: what a user of the reference would measure
        g = torch.Generator().manual_seed(0)
        photo = (want["rgb_fine"] + torch.randn(want["rgb_fine"].shape, generator=g) * 10 ** (-30 / 20)).clamp(0, 1)
        assert abs(helpers.psnr(got["rgb_fine"], photo) - helpers.psnr(want["rgb_fine"], photo)) < 0.05


def test_edit_shape_render_matches_oracle_directly(scene):
    """BASELINE configs[4] shape (test/config/edit_scannet_0113.yaml): 3 ray sets, ids [0, 4, 4], chunk 4096, on the bench
    scene, against the oracle port of render_rays_multi (pinned to the reference by the multi_* fixtures)."""
    from object_nerf_b200 import synthetic as S
    from object_nerf_b200.multi_rendering import render_rays_multi
    from oracle import onerf_oracle as O
    sc = bench.build_scene()
    n = 4096
    sel = torch.linspace(0, bench.N_RAYS - 1, n).long()
    rng = np.random.default_rng(11)
    sets = [sc["rays"][sel].clone()]
    for k in range(2):
        r = sets[0].clone()
        near = torch.from_numpy(rng.uniform(0.4, 1.2, size=n).astype(np.float32))
        far = near + torch.from_numpy(rng.uniform(0.2, 0.9, size=n).astype(np.float32))
        miss = torch.from_numpy(rng.random(n) < 0.3)
        near[miss] = 0
        far[miss] = 0
        r[:, 6], r[:, 7] = near, far
        sets.append(r)
    g = sc["grid"]
    grid = O.VoxelGrid(g["offset"], g["voxel_size"], g["shape"].tolist(), g["idx_map"], g["table"])
    with torch.no_grad():
        want = O.render_rays_multi(sc["weights"], grid, sc["code_table"], sets, [0, 4, 4], n_samples=64, n_importance=64)
    lib = S.make_code_library(sc["code_table"]).to(DEV)
    for precision, tol in (("fp32", 3e-4), ("bf16", 2e-2)):
        got = render_rays_multi(scene["models"], scene["embeddings"], lib, [s.to(DEV) for s in sets], [0, 4, 4], N_samples=64,
                                N_importance=64, chunk=4096, precision=precision)
        for k in ("rgb_fine", "depth_fine", "opacity_fine", "rgb_coarse"):
            err = (got[k].cpu() - want[k]).abs()
            assert (err > tol).float().mean().item() < 1e-2, (precision, k, err.max().item())
