"""Drop-in for the reference's `models/rendering.py` call surface: render_rays(), inference_model(),
sample_pdf() with identical signatures and result keys (reference models/rendering.py:11-17, 64-83,
233-250).  Every stage runs as a CUDA kernel of libonerf_sm100.so; there is no PyTorch or CPU
implementation behind these functions.

Extra keyword-only knobs (absorbed by **dummy_kwargs in the reference, so call sites stay valid):
  precision="bf16"|"fp32"   arithmetic of the fused encode+MLP kernel (default: env ONERF_PRECISION or bf16)
  _rand=dict(...)           test hook: pre-drawn random buffers (jitter, u, noise_*), see tests/synth.py
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from . import engine

__all__ = ["render_rays", "sample_pdf", "inference_model"]


def _is_voxel(embedding_xyz) -> bool:
    return hasattr(embedding_xyz, "voxel_idx_map")


def _grid_of(embedding_xyz):
    return engine.GridBuffers.from_module(embedding_xyz) if _is_voxel(embedding_xyz) else None


def _needs_grad(model, *tensors) -> bool:
    if not torch.is_grad_enabled():
        return False
    return any(p.requires_grad for p in model.parameters()) or any(
        t is not None and t.requires_grad for t in tensors)


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5, _u=None):
    """Reference models/rendering.py:11-61: draw N_importance depths from the piecewise-constant pdf
    `weights` (N, M) over `bins` (N, M+1).  eps is fixed at the reference default 1e-5 in the kernel."""
    if eps != 1e-5:
        raise RuntimeError("sample_pdf kernel is built with eps = 1e-5")
    seed = 0 if (det or _u is not None) else engine.new_seed()
    return engine.sample_pdf(bins, weights, N_importance, det, u=_u, seed=seed)


def inference_model(results: Dict[str, Any], model, embeddings: Dict[str, Any], typ: str, xyz, rays_d, z_vals,
                    chunk: int, noise_std: float, white_back: bool, is_eval: bool = False,
                    use_zero_as_last_delta: bool = False, forward_instance: bool = True,
                    embedding_instance: Optional[torch.Tensor] = None, frustum_bound_th: float = 0,
                    pass_through_mask: Optional[torch.Tensor] = None, rays_in_bbox: bool = False,
                    precision: Optional[str] = None, _rand: Optional[dict] = None, _rays=None, _seed=None,
                    **dummy_kwargs):
    """Encode + two-branch MLP + compositing for one pass; fills `results` in place with the reference's
    keys (models/rendering.py:64-230).  `chunk` is accepted and ignored: the kernels tile internally."""
    if _needs_grad(model, embedding_instance):
        raise NotImplementedError("inference_model() is the no-grad call surface; gradients flow through "
                                  "render_rays() (backward.RenderRaysFn)")
    n, s = z_vals.shape
    emb_xyz = embeddings["xyz"]
    use_voxel = _is_voxel(emb_xyz)
    packed = engine.packed_for(model, use_voxel)
    grid = _grid_of(emb_xyz)
    if _rays is None:  # explicit-xyz call surface: only the direction columns of `rays` are used
        _rays = torch.zeros(n, 8, dtype=torch.float32, device=z_vals.device)
        _rays[:, 3:6] = rays_d.reshape(n, 3)
        xyz_arg = xyz
    else:
        xyz_arg = None
    z = z_vals.contiguous()
    scene, obj = engine.field(_rays, z, packed, grid, codes=embedding_instance if forward_instance else None,
                              want_scene=True, want_object=forward_instance, precision=precision, xyz=xyz_arg)
    rand = _rand or {}
    seed = _seed if _seed is not None else (engine.new_seed() if noise_std > 0 else 0)
    out = engine.composite(z, scene, obj, noise_std=noise_std, white_back=white_back, is_eval=is_eval,
                           zero_last_delta=use_zero_as_last_delta, rays_in_bbox=rays_in_bbox,
                           frustum_bound_th=frustum_bound_th, pass_through_mask=pass_through_mask,
                           noise_scene=rand.get(f"noise_scene_{typ}"), noise_obj=rand.get(f"noise_obj_{typ}"),
                           seed=seed)
    results[f"weights_{typ}"] = out["weights"]
    results[f"opacity_{typ}"] = out["opacity"]
    results[f"z_vals_{typ}"] = z_vals
    results[f"rgb_{typ}"] = out["rgb"]
    results[f"depth_{typ}"] = out["depth"]
    if forward_instance:
        results[f"rgb_instance_{typ}"] = out["rgb_instance"]
        results[f"depth_instance_{typ}"] = out["depth_instance"]
        results[f"opacity_instance_{typ}"] = out["opacity_instance"]
    return



def query_sigma(model, embedding_xyz, xyz: torch.Tensor, obj_code: Optional[torch.Tensor] = None,
                chunk: int = 1 << 22, precision: Optional[str] = None) -> torch.Tensor:
    """Density of the field at arbitrary points: the `sigma_only=True` consumers of the fused encode + MLP
    (SURVEY.md section 8f row 4).  Replaces, in one call per chunk, the loop body of tools/extract_mesh.py:83-109
    (`embedding_xyz(xyz)` then `nerf_fine.forward(..., sigma_only=True)["sigma"]`, or with `obj_id > 0`
    `forward_instance(..., sigma_only=True)["inst_sigma"]` fed by one code-library row) and the density query of
    EmbeddingVoxel.self_pruning_empty_voxels (models/embedding_helper.py:219-225).
    xyz (B,3); obj_code None -> scene sigma (models/nerf_model.py:108-112), else (64,) -> object sigma (:140-144).
    Returns raw sigma (B,) (no relu), fp32, on xyz's device."""
    use_voxel = _is_voxel(embedding_xyz)
    packed = engine.packed_for(model, use_voxel)
    grid = _grid_of(embedding_xyz)
    xyz = xyz.reshape(-1, 3).contiguous().float()
    out = torch.empty(xyz.shape[0], dtype=torch.float32, device=xyz.device)
    for i in range(0, xyz.shape[0], chunk):
        pts = xyz[i:i + chunk]
        n = pts.shape[0]
        rays = torch.zeros(n, 8, dtype=torch.float32, device=xyz.device)   # directions are irrelevant for sigma
        z = torch.zeros(n, 1, dtype=torch.float32, device=xyz.device)
        scene, obj = engine.field(rays, z, packed, grid, code_row=obj_code, want_scene=obj_code is None,
                                  want_object=obj_code is not None, precision=precision, xyz=pts.view(n, 1, 3))
        out[i:i + n] = (scene if obj_code is None else obj)[:, 0, 3]
    return out

def _render_forward(cfg, rays, codes, keep=False):
    """The whole forward of render_rays on the CUDA kernels.  keep=True also returns what the backward needs."""
    rand = cfg["rand"]
    perturb, noise_std = cfg["perturb"], cfg["noise_std"]
    fi = cfg["forward_instance"]
    emb_xyz = cfg["embeddings"]["xyz"]
    use_voxel = _is_voxel(emb_xyz)
    grid = _grid_of(emb_xyz)
    seed = engine.new_seed() if (perturb > 0 or noise_std > 0) else 0
    z = engine.sample_coarse(rays, cfg["N_samples"], cfg["use_disp"], perturb, rand.get("jitter"), seed)
    results: Dict[str, Any] = {}
    saved: Dict[str, Any] = {}

    def one_pass(typ, z_vals, seed_off):
        model = cfg["models"][typ]
        packed = engine.packed_for(model, use_voxel, fresh=cfg.get("fresh_pack"))
        scene, obj = engine.field(rays, z_vals, packed, grid, codes=codes if fi else None, want_scene=True,
                                  want_object=fi, precision=cfg["precision"])
        ns, no = rand.get(f"noise_scene_{typ}"), rand.get(f"noise_obj_{typ}")
        if keep and noise_std > 0:      # the backward replays the noise: draw it into buffers instead of in-kernel
            ns = ns if ns is not None else torch.randn_like(z_vals)
            no = no if (no is not None or not fi) else torch.randn_like(z_vals)
        out = engine.composite(z_vals, scene, obj, noise_std=noise_std, white_back=cfg["white_back"],
                               is_eval=cfg["is_eval"], zero_last_delta=cfg["zero_last_delta"],
                               rays_in_bbox=cfg["rays_in_bbox"], frustum_bound_th=cfg["frustum_bound_th"],
                               pass_through_mask=cfg["pass_through_mask"], noise_scene=ns, noise_obj=no,
                               seed=seed + seed_off)
        results[f"weights_{typ}"] = out["weights"]
        results[f"opacity_{typ}"] = out["opacity"]
        results[f"z_vals_{typ}"] = z_vals
        results[f"rgb_{typ}"] = out["rgb"]
        results[f"depth_{typ}"] = out["depth"]
        if fi:
            results[f"rgb_instance_{typ}"] = out["rgb_instance"]
            results[f"depth_instance_{typ}"] = out["depth_instance"]
            results[f"opacity_instance_{typ}"] = out["opacity_instance"]
        if keep:
            saved[typ] = dict(z=z_vals, scene=scene, obj=obj, depth=out["depth"],
                              noise_scene=engine._f32(ns) if ns is not None else None,
                              noise_obj=engine._f32(no) if no is not None else None)

    one_pass("coarse", z, 1)
    if cfg["N_importance"] > 0:
        z_fine = engine.sample_pdf_merge(z, results["weights_coarse"], cfg["N_importance"], det=(perturb == 0),
                                         u=rand.get("u"), seed=seed + 2)
        one_pass("fine", z_fine, 3)
    return results, saved


def render_rays(models: Dict[str, Any], embeddings: Dict[str, Any], rays: torch.Tensor, N_samples: int = 64,
                use_disp: bool = False, perturb: float = 0, noise_std: float = 1, N_importance: int = 0,
                chunk: int = 1024 * 32, white_back: bool = False, forward_instance: bool = True,
                embedding_instance: Optional[torch.Tensor] = None, frustum_bound_th: float = 0,
                pass_through_mask: Optional[torch.Tensor] = None, rays_in_bbox: bool = False,
                **dummy_kwargs):
    """Reference models/rendering.py:233-337: stratified sampling -> coarse pass -> importance resampling
    -> fine pass.  rays (N,8) = [o, d, near, far]; returns the reference's result dict.  When parameters (or the
    object codes) require grad under torch.enable_grad(), the result carries autograd through backward.py."""
    rays = rays.contiguous().float()
    emb_xyz = embeddings["xyz"]
    cfg = dict(models=models, embeddings=embeddings, N_samples=N_samples, use_disp=use_disp, perturb=float(perturb),
               noise_std=float(noise_std), N_importance=N_importance, white_back=white_back,
               forward_instance=forward_instance, frustum_bound_th=float(frustum_bound_th),
               pass_through_mask=pass_through_mask, rays_in_bbox=rays_in_bbox,
               is_eval=bool(dummy_kwargs.get("is_eval", False)),
               zero_last_delta=bool(dummy_kwargs.get("use_zero_as_last_delta", False)),
               precision=dummy_kwargs.get("precision"), rand=dummy_kwargs.get("_rand") or {})
    codes = embedding_instance
    model_order = ["coarse"] + (["fine"] if N_importance > 0 else [])
    trainable = [p for typ in model_order for p in models[typ].parameters()]
    has_table = _is_voxel(emb_xyz)
    needs_grad = torch.is_grad_enabled() and (
        any(p.requires_grad for p in trainable) or (codes is not None and codes.requires_grad)
        or (has_table and emb_xyz.embedding_space_ftr.weight.requires_grad))
    if not needs_grad:
        return _render_forward(cfg, rays, codes)[0]
    # Training.  rays_in_bbox only swaps which weights feed the (detached) importance sampling and the returned
    # weights_* (models/rendering.py:228-229, :307): the gradients are unaffected.
    from . import backward
    cfg["has_table"], cfg["model_order"] = has_table, model_order
    params = ([emb_xyz.embedding_space_ftr.weight] if has_table else [])
    for typ in model_order:
        for w, b in engine.model_linears(models[typ]):
            params += [w, b]
    precision = cfg["precision"] or engine.default_precision()
    if precision == "bf16" and has_table:
        fn = backward.RenderRaysTcFn      # tcgen05 forward + backward
    else:
        # verification arithmetic (and the plain-PE model): fp32 forward AND backward, one function end to end
        cfg["precision"] = "fp32"
        cfg["fresh_pack"] = True          # training: never trust a cached blob (optimizers may write through .data)
        fn = backward.RenderRaysFn
    tensors = fn.apply(cfg, rays, codes, *params)
    keys = sorted(_result_keys(model_order, forward_instance))
    return dict(zip(keys, tensors))


def _result_keys(model_order, forward_instance):
    base = ["weights", "opacity", "z_vals", "rgb", "depth"] + (
        ["rgb_instance", "depth_instance", "opacity_instance"] if forward_instance else [])
    return [f"{k}_{typ}" for typ in model_order for k in base]
