"""In-kernel timeline of the two-tile forward kernel (build with `make -C object_nerf_b200/csrc TIMELINE=1`):
clock64() stamps of block 0, second tile pair.  Prints per slot: MMA warp (top, waits done, issue end) and epilogue
(start, end), and the encode warps' XS regenerations, in cycles relative to the first stamp."""
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
engine.field(rays, z, packed, grid, codes=codes, precision="bf16")
lib.onerf_debug_timeline2(buf.data_ptr())
engine.field(rays, z, packed, grid, codes=codes, precision="bf16")
torch.cuda.synchronize()
t = buf.cpu().tolist()
t0 = min(v for v in t if v > 10**9)
names = ["S0", "S1", "S2", "S3", "S4", "S5", "S6", "S7", "FIN", "DIR", "O0", "O1", "O2", "O3", "OFIN", "ODIR"]
# slot order: per layer: two-half -> (A,0)(A,1)(B,0)(B,1); one-half -> (A)(B)
slots = []
for l, nm in enumerate(names):
    two = l < 9
    for tile in "AB":
        for h in ((0, 1) if two else (0,)):
            slots.append(f"{tile}.{nm}.h{h}")
print(f"{'slot':12s} {'mma_top':>8s} {'waits':>8s} {'issued':>8s} {'w_full':>7s} | {'epi_start':>9s} {'acc_rdy':>8s} {'epi_end':>8s} {'work':>6s}  (cycles)")
for si, nm in enumerate(slots):
    r = [t[si * 8 + k] - t0 if t[si * 8 + k] > 0 else -1 for k in (0, 1, 2, 4, 6, 5)]
    print(f"{nm:12s} {r[0]:8d} {r[1]:8d} {r[2]:8d} {t[si * 8 + 3]:7d} | {r[3]:9d} {r[4]:8d} {r[5]:8d} {r[5] - r[4]:6d}")
print("slot 4 (A.S1.h0) per ring stage: before full wait, after wait, MMAs issued, after commit + syncwarp")
for gi in range(3):
    print("  stage", gi, [t[600 + gi * 4 + k] - t0 if t[600 + gi * 4 + k] > 0 else -1 for k in range(4)])
print("XS regenerations (before wait, after wait, done):")
for i in range(8):
    r = [t[512 + i * 4 + k] - t0 if t[512 + i * 4 + k] > 0 else -1 for k in range(3)]
    print(f"  use {i // 2} tile {'AB'[i % 2]}: {r}")
