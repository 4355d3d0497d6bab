"""bench.py — headline metric of BASELINE.json: rays/s rendering 640x480 frames (64 coarse + 64 importance
samples => 128-sample fine pass, scene + object branch, voxel embedding) through the reference call surface
`render_rays()`, plus the fused-MLP tensor-core roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

One process per GPU (torchrun for N > 1).  A step = one full frame (307 200 rays) per rank, rendered in
65 536-ray chunks; N > 1 is ray/tile-sharded inference: every rank renders its own frame and the tiles
(rgb, depth) are all-gathered over NCCL inside the timed region (weak scaling, no other collective).
Prints ONE JSON line (rank 0).  See DESIGN.md §measurement for the definitions.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W = 480, 640
N_RAYS = H * W
N_SAMPLES, N_IMPORTANCE = 64, 64
CHUNK = 65536
FLOP_PER_SAMPLE = 1776128          # 2 x MAC of the reference's nn.Linear layers, voxel config
FLOP_PER_RAY = (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) * FLOP_PER_SAMPLE
WORKLOAD = "configs[1]: 640x480 frame, 64 coarse + 64 importance (128-sample fine pass), scene+object two-branch, voxel embedding, eval"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.tmp = index, None, None

    def __enter__(self):
        try:
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.tmp is None:
            return out
        try:
            self.tmp.flush()
            rows = [r.strip().split(",") for r in open(self.tmp.name) if r.strip()]
            os.unlink(self.tmp.name)
            sm = [float(r[0]) for r in rows if len(r) >= 6]
            if sm:
                out["sm_mhz"] = statistics.median(sm)
                out["sm_max_mhz"] = float(rows[0][1])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for i, n in enumerate(names):
                    if any(r[2 + i].strip().lower().startswith("active") for r in rows if len(r) >= 6):
                        out["reasons"].append(n)
                out["samples"] = len(sm)
        except Exception:
            pass
        return out


def build_scene(device):
    from tests import helpers, synth
    wc = synth.make_weights(0, True, sigma_gain=8.0, sigma_bias=1.0)
    wf = synth.make_weights(1000, True, sigma_gain=8.0, sigma_bias=1.0)
    grid = synth.make_grid(seed=5, shape=(42, 42, 22), occupancy=0.6, voxel_size=0.05, n_rows=800000)
    rays = synth.pinhole_rays(H, W)                      # (307200, 8), pinhole 640x480, unit directions
    codes = synth.make_codes(2)
    ids = torch.from_numpy(__import__("numpy").random.default_rng(3).choice([4, 6], size=N_RAYS))
    return {"weights": {"coarse": wc, "fine": wf}, "grid": grid, "rays": rays, "codes": codes[ids]}


def run_reference(args, rank, world):
    """The reference's algorithm on the host cores: the oracle port (a torch-CPU restatement pinned bit-exactly
    to the reference, oracle/onerf_oracle.py), all host threads, a bounded ray sample per step."""
    if rank != 0:
        return
    from oracle import onerf_oracle as O
    cores = os.cpu_count() or 1
    sc = build_scene("cpu")
    g = sc["grid"]
    grid = O.VoxelGrid(g["offset"], g["voxel_size"], g["shape"].tolist(), g["idx_map"], g["table"])
    n = 2048
    sel = torch.linspace(0, N_RAYS - 1, n).long()
    rays, codes = sc["rays"][sel], sc["codes"][sel]
    with torch.no_grad():
        threads = pick_cpu_threads(lambda: O.render_rays(sc["weights"], grid, rays[:256], codes[:256], n_samples=N_SAMPLES,
                                                         n_importance=N_IMPORTANCE, is_eval=True))

    def step():
        with torch.no_grad():
            return O.render_rays(sc["weights"], grid, rays, codes, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE,
                                 is_eval=True)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = (f"{n} rays of the frame (every {N_RAYS // n}th) per step, oracle port (torch CPU fp32), {threads} torch "
              f"threads (fastest of a probe; host has {cores} logical cores)")
    print(json.dumps({
        "impl": "reference", "metric": "rays/s", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": n},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def pick_cpu_threads(fn):
    """torch's CPU matmuls on these small (32768 x 256) chunks get slower when oversubscribed (128 threads on
    the B200 host: 101 rays/s vs 468 rays/s at 16, profiles/r01_cpu_threads.md), so "all the host threads it
    can use" is found by timing a small probe at a few thread counts and keeping the fastest."""
    cores = os.cpu_count() or 1
    best, best_t = cores, None
    for t in sorted({cores, max(1, cores // 2), max(1, cores // 4), 32, 16}, reverse=True):
        if t > cores:
            continue
        torch.set_num_threads(t)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = t, dt
    torch.set_num_threads(best)
    return best


def cpu_baseline_sample():
    from oracle import onerf_oracle as O
    cores = os.cpu_count() or 1
    sc = build_scene("cpu")
    g = sc["grid"]
    grid = O.VoxelGrid(g["offset"], g["voxel_size"], g["shape"].tolist(), g["idx_map"], g["table"])
    n = 4096
    sel = torch.linspace(0, N_RAYS - 1, n).long()
    rays, codes = sc["rays"][sel], sc["codes"][sel]
    with torch.no_grad():
        threads = pick_cpu_threads(lambda: O.render_rays(sc["weights"], grid, rays[:256], codes[:256], n_samples=N_SAMPLES,
                                                         n_importance=N_IMPORTANCE, is_eval=True))
        t0 = time.perf_counter()
        reps = 0
        while reps < 2 or (time.perf_counter() - t0 < 10.0 and reps < 8):
            O.render_rays(sc["weights"], grid, rays, codes, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE, is_eval=True)
            reps += 1
        dt = time.perf_counter() - t0
    return {"value": n * reps / dt, "unit": "rays/s", "cores": threads, "kind": "port",
            "sample": f"{reps} x {n} rays of the frame, oracle port (torch CPU fp32), {threads} torch threads "
                      f"(fastest of a probe over thread counts; host has {cores} logical cores)"}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from object_nerf_b200 import Embedding, _lib, engine, render_rays
    from tests import helpers

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    sc = build_scene(dev)
    models = {k: helpers.make_model(w, True, dev) for k, w in sc["weights"].items()}
    emb = helpers.GridModule(sc["grid"]).to(dev)
    embeddings = {"xyz": emb, "dir": Embedding(3, 4)}
    rays_dev, codes_dev = sc["rays"].to(dev), sc["codes"].to(dev)
    rays_host, codes_host = sc["rays"].pin_memory(), sc["codes"].pin_memory()
    out_host = torch.empty(N_RAYS, 4, dtype=torch.float32).pin_memory()
    tiles = [torch.empty(N_RAYS, 4, device=dev) for _ in range(world)] if world > 1 else None

    def render(rays, codes):
        rgbd = torch.empty(N_RAYS, 4, device=dev)
        with torch.no_grad():
            for i in range(0, N_RAYS, CHUNK):
                r = render_rays(models, embeddings, rays[i:i + CHUNK], N_samples=N_SAMPLES, use_disp=False, perturb=0,
                                noise_std=0, N_importance=N_IMPORTANCE, chunk=32768, white_back=False,
                                embedding_instance=codes[i:i + CHUNK], is_eval=True, precision=args.precision)
                rgbd[i:i + CHUNK, :3] = r["rgb_fine"]
                rgbd[i:i + CHUNK, 3] = r["depth_fine"]
        return rgbd

    def step_device():
        rgbd = render(rays_dev, codes_dev)
        if world > 1:
            dist.all_gather(tiles, rgbd)
        return rgbd

    def step_e2e():
        r = rays_host.to(dev, non_blocking=True)
        c = codes_host.to(dev, non_blocking=True)
        rgbd = render(r, c)
        if world > 1:
            dist.all_gather(tiles, rgbd)
        out_host.copy_(rgbd, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        step_device()
    launches0 = _lib.launch_count(dev)
    with ClockSampler(local_rank) as cs:
        ms = timed(step_device, args.steps)
    launches = _lib.launch_count(dev) - launches0
    clocks = cs.summary()
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- roofline of the dominant kernel: the fine-pass field kernel (65 536 rays x 128 samples) ----
    engine.PROFILE_EVENTS = []
    step_device()
    torch.cuda.synchronize()
    fine = [a.elapsed_time(b) for (a, b, n, s) in engine.PROFILE_EVENTS if s == N_SAMPLES + N_IMPORTANCE and n == CHUNK]
    coarse = [a.elapsed_time(b) for (a, b, n, s) in engine.PROFILE_EVENTS if s == N_SAMPLES and n == CHUNK]
    field_ms_total = sum(a.elapsed_time(b) for (a, b, n, s) in engine.PROFILE_EVENTS)
    engine.PROFILE_EVENTS = None
    peaks, peaks_kind = load_peaks()
    fine_ms = statistics.mean(fine)
    flops = CHUNK * (N_SAMPLES + N_IMPORTANCE) * FLOP_PER_SAMPLE
    achieved = flops / (fine_ms * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"] if args.precision == "bf16" else 75.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "field_tc_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")

    if rank != 0:
        return
    value = N_RAYS * world * args.steps / (ms * 1e-3)
    e2e = N_RAYS * world * args.steps / (ms_e2e * 1e-3)
    cpu = cpu_baseline_sample() if (world == 1 and not args.no_cpu) else None
    line = {
        "metric": "rays/s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": N_RAYS, "chunk_rays": CHUNK,
                   "parallelism": f"ray-sharded dp{world}" if world > 1 else "single GPU",
                   "l2": "no explicit flush: each step streams ~3 GB of intermediates (>> 126 MB L2)",
                   "tflops_algorithmic_whole_step": N_RAYS * world * FLOP_PER_RAY * args.steps / (ms * 1e-3) / 1e12},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": N_RAYS * (8 + 64) * 4,
                "d2h_bytes_per_step": N_RAYS * 4 * 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "field_tc_kernel<voxel> fine pass (65536 rays x 128 samples)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "peak_source": f"{peaks_kind} bf16_tflops_sustained (kernel timed inside the step)",
                     "flops_per_launch": flops, "ms_per_launch": fine_ms,
                     "ms_per_launch_coarse": statistics.mean(coarse) if coarse else None,
                     "field_kernel_share_of_step": field_ms_total / (ms / args.steps),
                     "traffic": traffic},
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (development runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
