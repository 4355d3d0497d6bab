"""Time the oracle port at several torch thread counts (to pick an honest CPU baseline configuration)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oracle import onerf_oracle as O
sc = bench.build_scene("cpu")
g = sc["grid"]
grid = O.VoxelGrid(g["offset"], g["voxel_size"], g["shape"].tolist(), g["idx_map"], g["table"])
n = 1024
sel = torch.linspace(0, bench.N_RAYS - 1, n).long()
rays, codes = sc["rays"][sel], sc["codes"][sel]
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for t in (128, 64, 32, 16, 8):
    torch.set_num_threads(t)
    with torch.no_grad():
        O.render_rays(sc["weights"], grid, rays[:256], codes[:256], n_samples=64, n_importance=64, is_eval=True)
        t0 = time.perf_counter()
        O.render_rays(sc["weights"], grid, rays, codes, n_samples=64, n_importance=64, is_eval=True)
        dt = time.perf_counter() - t0
    print(f"threads {t}: {n/dt:.1f} rays/s ({dt:.2f} s)", flush=True)
