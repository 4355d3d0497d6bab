"""Dev check of the two-tile forward kernel: bit-exactness against the one-tile kernel and timing (CUDA events)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from object_nerf_b200 import Embedding, engine, synthetic as S

mode = sys.argv[1]
dev = torch.device("cuda", 0)
sc = bench.build_scene()
model = S.make_model(sc["weights"]["fine"], True, dev)
emb = S.GridModule(sc["grid"]).to(dev)
grid = engine.GridBuffers.from_module(emb)
packed = engine.packed_for(model, True)
n = int(os.environ.get("NRAYS", 65536))
rays = sc["rays"][:n].to(dev)
codes = sc["codes"][:n].to(dev)
for Sn in (128, 64):
    z = engine.sample_coarse(rays, Sn)
    for kw in (dict(want_scene=True, want_object=True), dict(want_scene=True, want_object=False), dict(want_scene=False, want_object=True)):
        sc_o, ob_o = engine.field(rays, z, packed, grid, codes=codes, precision="bf16", **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            engine.field(rays, z, packed, grid, codes=codes, precision="bf16", **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        flop = n * Sn * ((1399808 if kw["want_scene"] else 0) + (376320 if kw["want_object"] else 0))
        tag = f"S{Sn}_{int(kw['want_scene'])}{int(kw['want_object'])}"
        print(mode, tag, f"{ms:.3f} ms {flop / ms / 1e9:.0f} TFLOP/s")
        torch.save({"scene": sc_o, "obj": ob_o}, f"/tmp/tc2_{mode}_{tag}.pt")
if mode == "two":
    for f in sorted(os.listdir("/tmp")):
        if f.startswith("tc2_two_"):
            a, b = torch.load("/tmp/" + f), torch.load("/tmp/" + f.replace("two", "one"))
            for k in ("scene", "obj"):
                if a[k] is not None:
                    same = torch.equal(a[k], b[k])
                    print(f, k, "bit-identical" if same else f"DIFF max {(a[k]-b[k]).abs().max().item():.3e} frac {(a[k]!=b[k]).float().mean().item():.4f} nan {torch.isnan(a[k]).sum().item()}")
