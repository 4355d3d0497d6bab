"""Timeline of the pair2 kernel (build with `make TIMELINE=1`, run with ONERF_TC_VARIANT=pair2)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ONERF_TC_VARIANT"] = "pair2"
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
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.onerf_debug_timeline2.argtypes = [ctypes.c_void_p]
kw = dict(N_samples=64, perturb=0, noise_std=0, embedding_instance=codes, is_eval=True)
with torch.no_grad():
    render_rays(models, {"xyz": emb, "dir": Embedding(3, 4)}, rays, N_importance=64, **kw)
    torch.cuda.synchronize()
    lib.onerf_debug_timeline2(buf.data_ptr())
    render_rays(models, {"xyz": emb, "dir": Embedding(3, 4)}, rays, N_importance=0, **kw)
    torch.cuda.synchronize()
    lib.onerf_debug_timeline2(None)
t = buf.cpu().tolist()
t = [x & 0xFFFFFFFF for x in t]
t0 = t[128]
def d(a, b): return (a - b) & 0xFFFFFFFF
names = ["S0","S1","S2","S3","S4","S5","S6","S7","SFIN","SDIR","O0","O1","O2","O3","OFIN","ODIR"]
print("encode both tiles:", d(t[129], t0))
print("epilogues: layer half | tile0 ready done (dur) | tile1 ready done (dur)")
for l in range(16):
    for h in range(2):
        v = [d(t[((l*2+h)*2+tt)*2+k], t0) for tt in range(2) for k in range(2)]
        print(f"{names[l]:5s} h{h} | {v[0]:7d} {v[1]:7d} ({v[1]-v[0]:5d}) | {v[2]:7d} {v[3]:7d} ({v[3]-v[2]:5d})")
print("MMA stages: weights-ready time | full wait (since previous stage end) | tile0 (wait+issue) | tile1 (wait+issue)")
prev = None
for si in range(138):
    a, b, c = [t[130+si*3+k] for k in range(3)]
    if a == 0: break
    fw = d(a, prev) if prev is not None else 0
    if fw > 1 << 31: fw = 0
    print(f"{si:3d}: {d(a,t0) if d(a,t0) < (1<<31) else -d(t0,a):7d}  {fw:5d}  {d(b,a):5d}  {d(c,b):5d}")
    prev = c
