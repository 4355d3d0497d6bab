import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from object_nerf_b200 import engine, synthetic as S
dev = torch.device("cuda", 0)
sc = bench.build_scene()
model = S.make_model(sc["weights"]["fine"], True, dev)
emb = S.GridModule(sc["grid"]).to(dev)
grid = engine.GridBuffers.from_module(emb)
packed = engine.packed_for(model, True)
n = int(os.environ.get("NRAYS", 1024))
rays, codes = sc["rays"][:n].to(dev), sc["codes"][:n].to(dev)
z = engine.sample_coarse(rays, 64)
ws, wo = int(sys.argv[1]), int(sys.argv[2])
a, b = engine.field(rays, z, packed, grid, codes=codes, precision="bf16", want_scene=bool(ws), want_object=bool(wo))
torch.cuda.synchronize()
print("ok", None if a is None else a.abs().mean().item(), None if b is None else b.abs().mean().item())
