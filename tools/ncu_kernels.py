"""Workload for the ncu evidence of the small (HBM-bound) kernels: one 65 536-ray eval chunk of the bench frame
(sampling, ray constants, compositing, pdf merge, ...), one edit chunk (composite_multi) and TRAIN_STEPS training steps of
2 048 rays (compositing backward, head backward, column sums, ray sums, scatter).  Run under
  ncu --set full --clock-control none -k regex:'^(?!.*field_tc)' ...
and summarise with tools/ncu_summary.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from object_nerf_b200 import Embedding, render_rays, synthetic as S
from object_nerf_b200.losses import TotalLoss
from object_nerf_b200.ray_utils import camera_rays

dev = torch.device("cuda", 0)
sc = bench.build_scene()
bench.build_scene.cache = sc
models = {k: S.make_model(w, True, dev) for k, w in sc["weights"].items()}
emb = {"xyz": S.GridModule(sc["grid"]).to(dev), "dir": Embedding(3, 4)}
lib = S.make_code_library(sc["code_table"]).to(dev)
rays, ids = sc["rays"][:65536].to(dev), sc["ids"][:65536].to(dev)
with torch.no_grad():
    codes = lib.lookup(ids)
    for _ in range(int(os.environ.get("EVAL_REPS", 1))):
        render_rays(models, emb, rays, N_samples=64, perturb=0, noise_std=0, N_importance=64, embedding_instance=codes,
                    is_eval=True, precision="bf16")
    try:
        camera_rays(480, 640, 500.0, torch.eye(4)[:3], 0.05, 6.0, 1.0, device=dev)
    except Exception as ex:
        print("generate_rays skipped:", ex)
torch.cuda.synchronize()
# one edit chunk through the multi entry (composite_multi)
try:
    import types
    a = types.SimpleNamespace(steps=1, warmup=0, precision="bf16")
    os.environ["ONERF_EDIT_CHUNKS"] = "1"
    bench.run_edit(a, dev, sc, models, emb)
except Exception as ex:
    print("edit chunk skipped:", type(ex).__name__, ex)
# training steps
for m in models.values():
    m.train()
n = 2048
b = bench.train_batches(1, 0)[0]
b = {k: v.to(dev) for k, v in b.items()}
loss_fn = TotalLoss(bench.LOSS_CONF)
params = [p for m in models.values() for p in m.parameters()] + list(lib.parameters())
opt = torch.optim.Adam(params, lr=1e-3, fused=True)
for _ in range(int(os.environ.get("TRAIN_STEPS", 1))):
    opt.zero_grad(set_to_none=True)
    c = lib({"instance_ids": b["instance_ids"]})["embedding_instance"]
    out = render_rays(models, emb, b["rays"], N_samples=64, perturb=1.0, noise_std=1.0, N_importance=64,
                      embedding_instance=c, frustum_bound_th=0.025, pass_through_mask=b["pass_through_mask"], is_eval=False,
                      precision="bf16")
    loss_fn(out, b)[0].backward()
    opt.step()
torch.cuda.synchronize()
print("done")
