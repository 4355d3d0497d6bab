"""One 65 536-ray chunk of the bench workload through render_rays(), for ncu captures:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
   ncu --set full --clock-control none --import-source on -k regex:field_tc -s 2 -c 2 -o gpurun_out/field_tc python tools/profile_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from object_nerf_b200 import Embedding, render_rays  # noqa: E402
from tests import helpers  # noqa: E402

dev = torch.device("cuda", 0)
sc = bench.build_scene(dev)
models = {k: helpers.make_model(w, True, dev) for k, w in sc["weights"].items()}
emb = helpers.GridModule(sc["grid"]).to(dev)
n = int(os.environ.get("PROFILE_RAYS", bench.CHUNK))
rays, codes = sc["rays"][:n].to(dev), sc["codes"][:n].to(dev)
reps = int(os.environ.get("PROFILE_REPS", 2))
with torch.no_grad():
    for _ in range(reps):
        out = render_rays(models, {"xyz": emb, "dir": Embedding(3, 4)}, rays, N_samples=64, perturb=0, noise_std=0,
                          N_importance=64, embedding_instance=codes, is_eval=True,
                          precision=os.environ.get("ONERF_PRECISION", "bf16"))
torch.cuda.synchronize()
print("rgb_fine mean", out["rgb_fine"].mean().item())
