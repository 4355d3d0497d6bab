"""Debug: per-layer clock64() timeline of one tile pair of the CTA-pair field kernel (field_tc3.cu; blocks 0/1, second tile)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ONERF_TC_VARIANT"] = "pair3"
import torch
import bench
from object_nerf_b200 import Embedding, render_rays, _lib
from tests import helpers
dev = torch.device("cuda", 0)
sc = bench.build_scene(dev)
models = {k: helpers.make_model(w, True, dev) for k, w in sc["weights"].items()}
emb = helpers.GridModule(sc["grid"]).to(dev)
n = 65536
rays, codes = sc["rays"][:n].to(dev), sc["codes"][:n].to(dev)
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.onerf_debug_timeline3.argtypes = [ctypes.c_void_p]
with torch.no_grad():
    render_rays(models, {"xyz": emb, "dir": Embedding(3, 4)}, rays, N_samples=64, perturb=0, noise_std=0, N_importance=64,
                embedding_instance=codes, is_eval=True)
    torch.cuda.synchronize()
    lib.onerf_debug_timeline3(buf.data_ptr())
    render_rays(models, {"xyz": emb, "dir": Embedding(3, 4)}, rays, N_samples=64, perturb=0, noise_std=0, N_importance=0,
                embedding_instance=codes, is_eval=True)
    torch.cuda.synchronize()
    lib.onerf_debug_timeline3(None)
t = [x & 0xFFFFFFFF for x in buf.cpu().tolist()]
t0 = t[201]
def d(a): 
    v = (a - t0) & 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v
names = ["S0","S1","S2","S3","S4","S5","S6","S7","SFIN","SDIR","O0","O1","O2","O3","OFIN","ODIR"]
print(f"encode: {d(t[200])} cycles")
print("epilogue (compute warp 0):  rank0 acc_ready_seen / arrive_done | rank1 acc_ready_seen / arrive_done")
for l in range(16):
    for h in range(2):
        c0, d0 = d(t[(l*2+h)*4+2]), d(t[(l*2+h)*4+3])
        c1, d1 = d(t[768+(l*2+h)*4+2]), d(t[768+(l*2+h)*4+3])
        print(f"{names[l]:5s} h{h} | {c0:7d} {d0:7d} ({d0-c0:5d}) | {c1:7d} {d1:7d} ({d1-c1:5d})")
print("MMA warp stages: start, after E/X waits (flags), after full wait, after issue")
FL = {1:"H",2:"FIRST",4:"WX",8:"WE0",16:"WE1",32:"ACC",64:"h1",128:"WF1"}
for i in range(80):
    if t[256+i*3]==0: break
    a,b,c=[d(t[256+i*3+k]) for k in range(3)]
    e = d(t[512+i]); fl = t[640+i]
    fs = "|".join(v for k,v in FL.items() if fl&k)
    print(f"{i:3d}: start {a:7d}  Ewait {e-a:6d}  fullwait {b-e:6d}  issue {c-b:6d}   {fs}")
