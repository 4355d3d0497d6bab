"""Debug: per-layer clock64() timeline of one tile of the tcgen05 field kernel (block 0, second tile)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
buf = torch.zeros(256, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.onerf_debug_timeline.argtypes = [ctypes.c_void_p]
with torch.no_grad():
    render_rays(models, {"xyz": emb, "dir": Embedding(3, 4)}, rays, N_samples=64, perturb=0, noise_std=0, N_importance=64,
                embedding_instance=codes, is_eval=True)
    torch.cuda.synchronize()
    lib.onerf_debug_timeline(buf.data_ptr())
    render_rays(models, {"xyz": emb, "dir": Embedding(3, 4)}, rays, N_samples=64, perturb=0, noise_std=0, N_importance=0,
                embedding_instance=codes, is_eval=True)
    torch.cuda.synchronize()
    lib.onerf_debug_timeline(None)
t = buf.cpu().tolist()
t0 = t[201]
names = ["S0","S1","S2","S3","S4","S5","S6","S7","SFIN","SDIR","O0","O1","O2","O3","OFIN","ODIR"]
print(f"encode: {t[200]-t0} cycles")
print("layer half | mma_first_issue  mma_last_commit | acc_ready_seen  epi_done | issue_span  ready->done(epi)  prev_done->issue")
prev_done = t[200]
for l in range(16):
    for h in range(2):
        a, b, c, d = [t[(l*2+h)*4+k] - t0 for k in range(4)]
        print(f"{names[l]:5s} h{h} | {a:8d} {b:8d} | {c:8d} {d:8d} | {b-a:6d} {d-c:6d} {c-b:6d}")
