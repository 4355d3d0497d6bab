"""Drop-in for the reference's `render_tools/multi_rendering.py`: render_rays_multi() with the same
signature and result keys (reference render_tools/multi_rendering.py:160-175), used unchanged by
EditableRenderer.scene_inference / render_edit (render_tools/editable_renderer.py:125-140, 272-287).

Per ray set (scene id 0 or an object id): coarse depths -> one-branch field kernel (scene branch for id 0
with the removed-object box mask evaluated on the device, object branch with a constant code row
otherwise; zero-length rays muted) -> joint stable depth sort + compositing across all sets -> per-set
importance resampling -> fine pass.  No host round trips inside (the reference's check_in_any_boxes
goes device -> numpy -> device per chunk, utils/bbox_utils.py:119-130,170).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import _lib, engine
from .rendering import _grid_of, _is_voxel


def boxes_to_tensor(background_skip_bbox, device) -> Optional[torch.Tensor]:
    """Fold each BBoxRayHelper's xyz -> box-frame transform and bounds (utils/bbox_utils.py:119-130,
    158-186 with bbox_enlarge = 0, as check_in_any_boxes is called from multi_rendering.py:240) into
    rows [A (9) | t (3) | lo (3) | hi (3)] with p_box = A p + t."""
    if not background_skip_bbox:
        return None
    rows = []
    for _, box in background_skip_bbox.items():
        sf = float(box.scale_factor)
        P = np.asarray(box.pose_avg, dtype=np.float64).reshape(4, 4)
        Ax = np.asarray(box.axis_align_mat, dtype=np.float64).reshape(4, 4)
        A = Ax[:3, :3] @ P[:3, :3] * sf
        t = Ax[:3, :3] @ P[:3, 3] + Ax[:3, 3]
        bounds = np.asarray(box.bbox_bounds, dtype=np.float64)
        rows.append(np.concatenate([A.reshape(-1), t, bounds[0], bounds[1]]))
    return torch.from_numpy(np.stack(rows).astype(np.float32)).to(device)


def render_rays_multi(models: Dict[str, Any], embeddings: Dict[str, Any], code_library, rays_list: list,
                      obj_instance_ids: list, N_samples: int = 64, use_disp: bool = False, perturb: float = 0,
                      noise_std: float = 0, N_importance: int = 0, chunk: int = 1024 * 32,
                      white_back: bool = False, background_skip_bbox: Dict[str, Any] = None,
                      precision: Optional[str] = None, _staged: bool = False):
    """Reference render_tools/multi_rendering.py:160-175.  The whole forward is ONE C call (onerf_render_multi_fwd);
    `_staged=True` runs the same kernels stage by stage from Python (tests: both routes are bit-identical)."""
    assert len(rays_list) == len(obj_instance_ids)
    if noise_std != 0:
        raise NotImplementedError("render_rays_multi kernels are built for noise_std = 0 "
                                  "(the only value EditableRenderer passes)")
    emb_xyz = embeddings["xyz"]
    if not _is_voxel(emb_xyz):
        raise RuntimeError("render_rays_multi requires the voxel embedding, as the reference does "
                           "(render_tools/multi_rendering.py:55 unpacks a tuple)")
    if any(r.shape[1] != 8 for r in rays_list):
        raise NotImplementedError("10-column (bbox-clipped) rays are not built (unused by the demo)")
    grid = _grid_of(emb_xyz)
    dev = rays_list[0].device
    n_obj, n, s = len(rays_list), rays_list[0].shape[0], N_samples
    rays_list = [r.contiguous().float() for r in rays_list]
    boxes = boxes_to_tensor(background_skip_bbox, dev)
    code_table = engine._f32(code_library.embedding_instance.weight.detach())
    if not _staged:
        return _render_multi_one_call(models, grid, code_table, rays_list, [int(i) for i in obj_instance_ids], N_samples,
                                      use_disp, perturb, N_importance, white_back, boxes, precision)

    def eval_pass(model, z_all):
        packed = engine.packed_for(model, True)
        s_ = z_all.shape[2]
        field_all = torch.empty(n_obj, n, s_, 4, dtype=torch.float32, device=dev)
        for i, iid in enumerate(obj_instance_ids):
            is_obj = iid > 0
            engine.field(rays_list[i], z_all[i], packed, grid, code_row=code_table[iid] if is_obj else None,
                         want_scene=not is_obj, want_object=is_obj, precision=precision, mute_zero_rays=True,
                         boxes=None if is_obj else boxes,
                         scene_out=None if is_obj else field_all[i], obj_out=field_all[i] if is_obj else None)
        return field_all

    results: Dict[str, Any] = {}
    with torch.no_grad():
        z_all = torch.empty(n_obj, n, s, dtype=torch.float32, device=dev)
        for i in range(n_obj):
            engine.sample_coarse(rays_list[i], s, use_disp, 0.0, out=z_all[i])
        out = engine.composite_multi(z_all, eval_pass(models["coarse"], z_all), white_back, want_ids=True,
                                     want_unsorted=N_importance > 0)
        for k in ("weights", "opacity", "z_vals", "rgb", "depth"):
            results[f"{k}_coarse"] = out[k]
        results["obj_ids_coarse"] = out["obj_ids"]
        if N_importance > 0:
            z_fine = torch.empty(n_obj, n, s + N_importance, dtype=torch.float32, device=dev)
            det = (perturb == 0)
            for i in range(n_obj):
                engine.sample_pdf_merge(z_all[i], out["weights_unsorted"][i], N_importance, det,
                                        seed=0 if det else engine.new_seed(), out=z_fine[i])
            out = engine.composite_multi(z_fine, eval_pass(models["fine"], z_fine), white_back)
            for k in ("weights", "opacity", "z_vals", "rgb", "depth"):
                results[f"{k}_fine"] = out[k]
    return results


def _render_multi_one_call(models, grid, code_table, rays_list, obj_ids, n_samples, use_disp, perturb, n_importance,
                           white_back, boxes, precision):
    lib = _lib.load()
    dev = rays_list[0].device
    n_obj, n = len(rays_list), rays_list[0].shape[0]
    f = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
    a = _lib.RenderMultiArgs()
    rays_p = (C.c_void_p * n_obj)(*[r.data_ptr() for r in rays_list])
    ids_p = (C.c_int * n_obj)(*obj_ids)
    a.rays_list_host, a.obj_ids_host = rays_p, ids_p
    a.n_obj, a.n_rays, a.n_samples, a.n_importance = n_obj, n, n_samples, n_importance
    a.grid = C.pointer(grid.c)
    packed_c = engine.packed_for(models["coarse"], True)
    packed_f = engine.packed_for(models["fine"], True) if n_importance > 0 else None
    a.packed_coarse = packed_c.data_ptr()
    a.packed_fine = packed_f.data_ptr() if packed_f is not None else None
    a.code_table, a.n_codes = code_table.data_ptr(), code_table.shape[0]
    a.precision = engine.PRECISIONS[precision or engine.default_precision()]
    a.use_disp, a.perturb = int(bool(use_disp)), float(perturb)
    a.seed = 0 if perturb == 0 else engine.new_seed()
    a.white_back = int(bool(white_back))
    a.boxes, a.n_boxes = _lib.ptr(boxes), (boxes.shape[0] if boxes is not None else 0)
    results: Dict[str, Any] = {}
    for typ, s in (("coarse", n_samples), ("fine", n_samples + n_importance)):
        if typ == "fine" and n_importance == 0:
            continue
        t = n_obj * s
        m = dict(weights=f(n, t), opacity=f(n), z_vals=f(n, t), rgb=f(n, 3), depth=f(n))
        if typ == "coarse":
            m["obj_ids"] = f(n, t)
        cm = getattr(a, typ)
        for k, v in m.items():
            setattr(cm, k, v.data_ptr())
            results[f"{k}_{typ}"] = v
    ws = torch.empty(max(lib.onerf_render_multi_workspace_bytes(n, n_obj, n_samples, n_importance), 256), dtype=torch.uint8,
                     device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    with torch.cuda.device(dev):
        _lib.check(lib.onerf_render_multi_fwd(_lib.ctx(dev), C.byref(a), _lib.stream()))
    return results
