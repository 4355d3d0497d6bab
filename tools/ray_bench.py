"""Timing of the fused camera-ray kernel (onerf_camera_rays: pixel grid + pose + box -> (N,8) rays), CUDA events.
Algorithmic bytes per ray: 32 B rays + 1 B mask written, nothing read."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from object_nerf_b200 import ray_utils  # noqa: E402
from tests import cases  # noqa: E402
from tests.test_gpu_rays import Box  # noqa: E402

dev = torch.device("cuda", 0)
box = Box(cases.build_bbox_case(cases.BBOX_CASES["bbox_basic"]))
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
for (h, w) in ((480, 640), (2160, 3840), (8192, 8192)):
    cam = cases.build_camera_case(dict(H=h, W=w, fovx_deg=70.0, seed=1))
    for _ in range(5):
        ray_utils.camera_rays(h, w, cam["focal"], cam["c2w"], 0.3, 6.0, 2.0, box=box, device=dev, return_mask=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ray_utils.camera_rays(h, w, cam["focal"], cam["c2w"], 0.3, 6.0, 2.0, box=box, device=dev, return_mask=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = h * w * 33 / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "camera_rays_kernel (+ torch.empty of the outputs)", "H": h, "W": w, "ms": ms, "rays_per_s": h * w / (ms * 1e-3),
                      "algorithmic_GBps": gbs, "hbm_peak_GBps": peaks.get("hbm_gbs")}))
