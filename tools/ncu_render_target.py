#!/usr/bin/env python
"""Two 400x600 frames of NeRFAll.forward in eval mode (the render leg of bench.py, same arguments) for a launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_render.csv \\
        python tools/ncu_render_target.py
Without ncu it prints the device time of the second frame and the time between the frame's first and last kernel that
no kernel of the list accounts for is host-side (launch gaps, synchronisations)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LUSH_VERBOSE", "0")
import bench  # noqa: E402
from lush_nerf_b200.synthetic import build_model, poster_args, spiral_poses  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = build_model(poster_args()).to(dev).eval()
scene = bench.Scene(0)
poses = torch.from_numpy(spiral_poses(4, radius_scale=2.0)).to(dev)
rk = dict(perturb=False, N_importance=bench.N_IMPORTANCE, N_samples=bench.N_SAMPLES, use_viewdirs=True, white_bkgd=False,
          raw_noise_std=0.0, inference=True, near=0.0, far=1.0)
with torch.no_grad():
    for i in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rgbs, _, _ = model(bench.H, bench.W, scene.K, chunk=1024 * 32, poses=poses[2 + i:3 + i], render_kwargs=dict(rk))
        b.record()
        torch.cuda.synchronize()
        print(f"frame {i}: {a.elapsed_time(b):.2f} ms")
