"""One launch of the two-tile field kernel for ncu (bench scene, 32768 rays x 128 samples, both branches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from object_nerf_b200 import engine, synthetic as S
dev = torch.device("cuda", 0)
sc = bench.build_scene()
model = S.make_model(sc["weights"]["fine"], True, dev)
grid = engine.GridBuffers.from_module(S.GridModule(sc["grid"]).to(dev))
packed = engine.packed_for(model, True)
n = int(os.environ.get("NRAYS", 32768))
rays, codes = sc["rays"][:n].to(dev), sc["codes"][:n].to(dev)
z = engine.sample_coarse(rays, 128)
for _ in range(2):
    engine.field(rays, z, packed, grid, codes=codes, precision="bf16")
torch.cuda.synchronize()
