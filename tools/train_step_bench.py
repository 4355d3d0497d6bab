"""configs[2] in the bench scene: one training step of 2048 rays (render_rays train mode -> TotalLoss -> backward -> Adam),
timed with CUDA events.  Reports ms per step and the split forward / backward."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from object_nerf_b200 import Embedding, render_rays
from tests import cases, helpers

dev = torch.device("cuda", 0)
sc = bench.build_scene(dev)
models = {k: helpers.make_model(w, True, dev).train() for k, w in sc["weights"].items()}
emb = helpers.GridModule(sc["grid"]).to(dev)
lib = helpers.CodeLib(__import__("tests.synth", fromlist=["x"]).make_codes(2)).to(dev)
n = int(os.environ.get("TRAIN_RAYS", 2048))
rng = np.random.default_rng(0)
sel = torch.from_numpy(rng.integers(0, bench.N_RAYS, size=n))
rays = sc["rays"][sel].to(dev)
ids = torch.from_numpy(rng.choice([4, 6], size=n)).to(dev)
batch = {"rgbs": torch.rand(n, 3, device=dev), "depths": torch.rand(n, device=dev) * 2 + 0.3,
         "valid_mask": torch.rand(n, device=dev) < 0.9, "instance_mask": torch.rand(n, device=dev) < 0.5,
         "instance_mask_weight": torch.where(torch.rand(n, device=dev) < 0.5, 1.0, 0.05)}
ptm = torch.rand(n, 1, device=dev) < 0.5
params = [p for m in models.values() for p in m.parameters()] + list(lib.parameters()) + list(emb.parameters())
opt = torch.optim.Adam(params, lr=1e-3, fused=True)
precision = os.environ.get("ONERF_PRECISION", "bf16")
from object_nerf_b200.losses import TotalLoss
loss_fn = TotalLoss({k: v for k, v in cases.LOSS_CONF.items()})
STEPS = int(os.environ.get("TRAIN_STEPS", 10))

def step():
    opt.zero_grad(set_to_none=True)
    codes = lib.embedding_instance(ids)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    out = render_rays(models, {"xyz": emb, "dir": Embedding(3, 4)}, rays, N_samples=64, perturb=1.0, noise_std=1.0,
                      N_importance=64, embedding_instance=codes, frustum_bound_th=0.025, pass_through_mask=ptm,
                      is_eval=False, precision=precision)
    loss, _ = loss_fn(out, batch)
    e[1].record()
    loss.backward()
    e[2].record()
    opt.step()
    e[3].record()
    torch.cuda.synchronize()
    return loss.item(), [e[i].elapsed_time(e[i + 1]) for i in range(3)]

for _ in range(3):
    step()
ts = [step() for _ in range(STEPS)]
fw = np.mean([t[1][0] for t in ts]); bw = np.mean([t[1][1] for t in ts]); ad = np.mean([t[1][2] for t in ts])
print(f"train step, {n} rays, forward precision {precision}: forward+loss {fw:.1f} ms, backward {bw:.1f} ms, adam {ad:.1f} ms, "
      f"total {fw+bw+ad:.1f} ms = {n/(fw+bw+ad)*1e3:.0f} rays/s; losses {[round(t[0],4) for t in ts]}")
