"""Where each role of the two-tile kernel waits (library built with -DONERF_WAITSTATS, see ab/): block 0, whole launch."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from object_nerf_b200 import _lib, engine, synthetic as S
dev = torch.device("cuda", 0)
sc = bench.build_scene()
model = S.make_model(sc["weights"]["fine"], True, dev)
emb = S.GridModule(sc["grid"]).to(dev)
grid = engine.GridBuffers.from_module(emb)
packed = engine.packed_for(model, True)
rays, codes = sc["rays"][:65536].to(dev), sc["codes"][:65536].to(dev)
z = engine.sample_coarse(rays, 128)
buf = torch.zeros(1024, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.onerf_debug_timeline2.argtypes = [ctypes.c_void_p]
kw = dict(want_scene=os.environ.get("BR", "11")[0] == "1", want_object=os.environ.get("BR", "11")[1] == "1")
engine.field(rays, z, packed, grid, codes=codes, precision="bf16", **kw)
lib.onerf_debug_timeline2(buf.data_ptr())
engine.field(rays, z, packed, grid, codes=codes, precision="bf16", **kw)
torch.cuda.synchronize()
t = buf.cpu().tolist()
NAMES = {
    "epilogue warp 0": ["ld+release", "math h0 (stash)", "math h1/one-half", "st+wait+arrive", "dir event", "xgen (incl. its waits)", "wait acc_ready", "-"],
    "producer": ["wait empty"] + ["-"] * 7,
    "mma": ["-", "wait acc_free", "wait h_ready", "wait xs_ready", "wait full", "issue", "-", "commit"],
    "encode 18": ["-"] * 5 + ["wait f_free", "-", "-"],
    "encode 19": ["-"] * 5 + ["wait f_free", "-", "-"],
}
n_events = None
for role, idx in (("epilogue warp 0", 0), ("producer", 1), ("mma", 2), ("encode 18", 3), ("encode 19", 4)):
    o = t[900 + idx * 10: 900 + idx * 10 + 10]
    tot = o[8]
    print(f"{role:16s} total {tot:>10d} cycles:", {n: f"{100 * v / max(tot, 1):.0f}%" for n, v in zip(NAMES[role], o[:8]) if v and n != "-"},
          "rest", f"{100 * (tot - sum(v for n, v in zip(NAMES[role], o[:8]) if n != '-')) / max(tot, 1):.0f}%")
