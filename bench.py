#!/usr/bin/env python
"""bench.py -- train rays/s of the LuSh-NeRF per-ray rendering hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling weak|strong]
                    [--mode train|stress] [--graph single|split|none]

mode train (default).  One "step" = one pass of the hot path over one batch of synthetic LLFF-shaped pixels at the
poster_lushnerf training configuration (BASELINE.json configs[1]): batch staging from (image, pixel) indices,
NeRFAll.forward in rigid-blur-kernel mode (N_rand=1024 pixels x K=5 warped rays = 5120 rays, 64 coarse + 64 fine
samples, perturb=1, raw_noise_std=1) + loss + backward + gradient all-reduce + Adam.  For N>1 every rank renders its
own N_rand pixels (--scaling weak, default) or its 1/N share of ONE global N_rand=1024 batch (--scaling strong, the
reference's DataParallel semantics, run_lushnerf.py:348 = BASELINE configs[2]); the only collective is the in-place
all-reduce of the flat gradient arenas (5.2 MB).

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with the batch already resident in HBM;
`e2e` = the same step driven from pinned host buffers (H2D of the batch indices + D2H of the loss inside the timed
region); `roofline` = algorithmic MLP FLOPs / measured MLP device time against the measured bf16 peak;
`cpu_baseline` = the CPU oracle port on this box's host cores on a bounded sample; `render` = ms/frame of
--render_only style full frames (BASELINE configs[3]; frames sharded over the ranks for N>1).
`--impl reference` times the reference algorithm's CPU implementation (oracle port; the reference tree itself cannot
travel to the GPU box) on the host cores for the same metric and workload.

mode stress (BASELINE configs[4]): 65 536-ray batches x S = N in {32, 64, 128, 256}: compositing forward / backward,
sample_pdf (det, strict and fused cdf) and the fused sample_merge, each against the measured HBM copy bandwidth with
SURVEY.md 8d's algorithmic bytes; inputs rotate through > 2x L2 of buffers so that no launch finds its input in L2.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work: scene MLP 593 408 MAC/sample, fwd+bwd = 3x forward
FLOP_FWD_PER_SAMPLE = 2 * 593408
N_RAND, K_RAYS, N_SAMPLES, N_IMPORTANCE = 1024, 5, 64, 64
SAMPLES_PER_RAY = N_SAMPLES + (N_SAMPLES + N_IMPORTANCE)
H, W, FOCAL = 400, 600, 500.0
VIEWS, LLFFHOLD = 20, 8
L2_BYTES = 126 * 1024 * 1024


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "tflops": d.get("bf16_tflops_sustained", 1400.0),
                "tflops_burst": d.get("bf16_tflops", 1590.0),
                "source": "MEASURED_PEAKS.json (bf16 sustained: kernel timed inside a long step)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def load_traffic():
    """dram bytes (read + write) of the scene-MLP kernels of ONE training step, from the newest committed ncu
    `--set full` capture of this very command (profiles/rNN_ncu_mlp_step.json, written by tools/ncu_step.py)."""
    for name in ("r02_ncu_mlp_step.json", "r01_ncu_mlp_step.json"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            with open(path) as f:
                d = json.load(f)
            return {"bytes_per_step": d["dram_bytes_per_step"], "source": f"profiles/{name}"}
        except Exception:
            continue
    return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=100):
        self.idx, self.proc, self.path, self.period = gpu_index, None, None, int(period_ms)

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", str(self.period)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------------
# synthetic scene: 20 views of 400x600, 17 training views (llffhold = 8)
# ------------------------------------------------------------------------------------------------
class Scene:
    """Device- or host-side tables of the synthetic LLFF-shaped scene and its index batches."""

    def __init__(self, seed=0):
        from lush_nerf_b200.synthetic import synthetic_poses
        self.poses = torch.from_numpy(synthetic_poses(VIEWS, seed))
        i_test = np.arange(VIEWS)[::LLFFHOLD]
        self.i_train = np.array([i for i in range(VIEWS) if i not in i_test])
        g = torch.Generator().manual_seed(seed)
        self.images = torch.rand(VIEWS, H, W, 3, generator=g)                 # targets ~ U[0,1]
        self.masks = torch.rand(VIEWS, H, W, generator=g) < 0.5               # fq_mask ~ Bernoulli(0.5)
        self.K = np.array([[FOCAL, 0, 0.5 * W], [0, FOCAL, 0.5 * H], [0, 0, 1]])

    def index_batch(self, n, gen):
        """n random (training view, pixel) pairs: the whole per-step host->device traffic (16 B per pixel)."""
        v = torch.from_numpy(self.i_train)[torch.randint(0, len(self.i_train), (n,), generator=gen)]
        p = torch.randint(0, H * W, (n,), generator=gen)
        return dict(images_idx=v.reshape(n, 1).contiguous(), pix_idx=p.contiguous())

    def expand(self, b):
        """The reference's batch form of an index batch (host tensors): rays [n,3,2], rgbsf, fq_mask, images_idx."""
        from lush_nerf_b200.utils.run_lushnerf_helpers import get_rays_np
        v, p = b["images_idx"].reshape(-1), b["pix_idx"]
        rays = torch.empty(len(v), 3, 2)
        for vi in torch.unique(v).tolist():
            ro, rd = get_rays_np(H, W, self.K, self.poses[vi].numpy())
            sel = (v == vi).nonzero().reshape(-1)
            px = p[sel].numpy()
            rays[sel, :, 0] = torch.from_numpy(np.ascontiguousarray(ro.reshape(-1, 3)[px])).float()
            rays[sel, :, 1] = torch.from_numpy(rd.reshape(-1, 3)[px].astype(np.float32))
        return dict(rays=rays, rgbsf=self.images.reshape(VIEWS, -1, 3)[v, p], fq_mask=self.masks.reshape(VIEWS, -1)[v, p],
                    images_idx=b["images_idx"])


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores, bounded sample of the SAME workload
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, budget_s=150.0, chunk_px=128):
    """fwd+bwd of the poster training step on the CPU oracle.  A step processes `n_px` pixels (x 5 rays x 64+64
    samples) of the stated workload in chunks of `chunk_px` pixels with gradient accumulation (the oracle keeps ~20 KB
    of fp32 activations per sample); n_px = 1024 (the full N_rand batch) unless that would exceed `budget_s` for the
    whole run, in which case the step is the largest multiple of chunk_px that fits (stated in `sample`).
    Returns (rays/s, ms/step, cores, sample)."""
    from oracle import lush_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nets = O.det_nets(O.POSTER_CFG, tag=1, sigma_gain=30.0)
    rbk = O.det_rbk_params(tag=7)
    pn = {k: {n: t.clone().requires_grad_(True) for n, t in d.items()} for k, d in nets.items()}
    pr = {n: t.clone().requires_grad_(True) for n, t in rbk.items()}
    scene = Scene(0)
    gen = torch.Generator().manual_seed(1)

    def chunk_pass(b, scale):
        R = b["rays"].shape[0] * K_RAYS
        out, _ = O.nerfall_train_forward(
            pn, pr, O.POSTER_CFG, b["rays"].contiguous(), b["images_idx"], H, W, FOCAL, N_samples=N_SAMPLES,
            N_importance=N_IMPORTANCE, allkernel=True, kernel_pixel=b["fq_mask"], t_rand=torch.rand(R, N_SAMPLES),
            noise_c=torch.randn(R, N_SAMPLES - 1), noise_f=torch.randn(R, N_SAMPLES + N_IMPORTANCE - 1),
            u=torch.rand(R, N_IMPORTANCE))
        (O.blur_loss(out[0], out[1], b["rgbsf"]) * scale).backward()

    probe = scene.expand(scene.index_batch(chunk_px, gen))
    chunk_pass(probe, 1.0)                                   # untimed: page-in, thread pools
    t0 = time.perf_counter()
    chunk_pass(probe, 1.0)
    t_chunk = time.perf_counter() - t0
    fit = int(budget_s / max(t_chunk * (steps + warmup), 1e-9))
    n_chunks = max(1, min(N_RAND // chunk_px, fit))
    n_px = n_chunks * chunk_px
    times = []
    for it in range(warmup + steps):
        batch = scene.expand(scene.index_batch(n_px, gen))
        for d in list(pn.values()) + [pr]:
            for t in d.values():
                t.grad = None
        t0 = time.perf_counter()
        for c in range(n_chunks):
            sl = slice(c * chunk_px, (c + 1) * chunk_px)
            chunk_pass({k: v[sl] for k, v in batch.items()}, 1.0 / n_chunks)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    R = n_px * K_RAYS
    sample = (f"{n_px} of the {N_RAND} pixels of a step x {K_RAYS} rays = {R} rays/step, {N_SAMPLES}+{N_IMPORTANCE} samples, "
              f"{VIEWS - len(np.arange(VIEWS)[::LLFFHOLD])} training views, NeRFAll.forward RBK mode fwd+bwd in {n_chunks} "
              f"chunk(s) of {chunk_px} px with gradient accumulation, {steps} timed steps, torch CPU fp32, {cores} threads")
    return R * len(times) / total, 1e3 * total / len(times), cores, sample


def workload_config(n_gpus, scaling="weak", graph=None):
    per_gpu = N_RAND if scaling == "weak" else N_RAND // max(n_gpus, 1)
    cfg = {"workload": "poster_lushnerf training step (BASELINE configs[1]): NeRFAll.forward RBK mode, "
                       "N_rand=1024 px x K=5 rays = 5120 rays/GPU, N_samples=64, N_importance=64, perturb=1, "
                       "raw_noise_std=1, scene+noise MLPs, loss, backward, Adam",
           "rays_per_step_per_gpu": per_gpu * K_RAYS, "image": f"{H}x{W} f={FOCAL}", "views": VIEWS,
           "train_views": VIEWS - len(np.arange(VIEWS)[::LLFFHOLD]),
           "batch": "random (training view, pixel) index pairs; rays / targets / masks staged on the device from the "
                    "resident poses / images / masks",
           "parallelism": f"rays sharded x{n_gpus}, in-place all-reduce of the gradient arenas" if n_gpus > 1 else "single GPU",
           "l2": "per-step working set (GBs of activation stash) exceeds the 126 MB L2; no flush needed"}
    if scaling == "strong":
        cfg["workload"] = ("poster_lushnerf training step, ONE global N_rand=1024 batch split over the GPUs "
                           "(BASELINE configs[2], DataParallel semantics): " + cfg["workload"].split(": ", 1)[1]
                           .replace("5120 rays/GPU", f"{per_gpu * K_RAYS} rays/GPU"))
    if graph is not None:
        cfg["cuda_graph"] = graph
    return cfg


def run_reference_arm(args, rank):
    if rank != 0:
        return
    val, ms, cores, sample = cpu_reference_run(args.steps, args.warmup)
    cfg = workload_config(1)
    cfg["cpu_sample"] = sample
    line = {"impl": "reference", "metric": "train rays/s (coarse+fine fwd+bwd, 64+64 samples)", "value": val,
            "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# render leg (BASELINE configs[3])
# ------------------------------------------------------------------------------------------------
def render_leg(args, model, K, ops, mlp_events, dev, local, rank, world, barrier):
    """ms/frame of NeRFAll.forward in eval mode over a spiral path, render_radius_scale 2.0 (rays generated from the
    pose on the device).  N > 1: the frames are sharded over the ranks the way the reference shards poses on dim 0
    (run_lushnerf.py:441-450, 492-496); ms/frame = max-over-ranks device time / total frames."""
    from lush_nerf_b200.synthetic import spiral_poses
    per_rank = max(1, args.render_frames // world)
    total = per_rank * world
    poses = torch.from_numpy(spiral_poses(total + 2, radius_scale=2.0)).to(dev)
    mine = poses[2 + rank * per_rank: 2 + (rank + 1) * per_rank]
    rk = dict(perturb=False, N_importance=N_IMPORTANCE, N_samples=N_SAMPLES, use_viewdirs=True, white_bkgd=False,
              raw_noise_std=0.0, inference=True, near=0.0, far=1.0)
    model.eval()
    with torch.no_grad():
        rsampler = ClockSampler(local, period_ms=50)
        if rank == 0:
            rsampler.start()
        model(H, W, K, chunk=1024 * 32, poses=poses[:2], render_kwargs=dict(rk))   # warm-up frames (sampler already on)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        mlp_events.clear()
        ops.MLP_EVENT_SINK = mlp_events
        try:
            r0.record()
            rgbs, _, _ = model(H, W, K, chunk=1024 * 32, poses=mine, render_kwargs=dict(rk))
            r1.record()
            barrier()
        finally:
            ops.MLP_EVENT_SINK = None
    rclocks = rsampler.stop() if rank == 0 else None
    ms = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
    finite = torch.tensor([1 if bool(torch.isfinite(rgbs).all().item()) else 0], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(finite, op=dist.ReduceOp.MIN)
    ms_frame = ms.item() / total
    mlp_ms_frame = sum(a.elapsed_time(b) for lab, a, b in mlp_events if lab == "fwd") / per_rank
    fl = FLOP_FWD_PER_SAMPLE * H * W * SAMPLES_PER_RAY
    return {"ms_per_frame": ms_frame, "frames": total, "frames_per_gpu": per_rank, "n_gpus": world,
            "rays_per_s": H * W / (ms_frame * 1e-3), "image": f"{H}x{W}", "chunk": 32768,
            "sampling": "det, strict cdf (reference op order): fine-sample indices bit-exact",
            "mlp_ms_per_frame_per_gpu": mlp_ms_frame,
            "mlp_tflops_per_gpu": fl / (mlp_ms_frame * 1e-3) / 1e12 if mlp_ms_frame > 0 else None,
            "finite": bool(finite.item()), "clocks": rclocks}


# ------------------------------------------------------------------------------------------------
# stress sweep (BASELINE configs[4])
# ------------------------------------------------------------------------------------------------
def stress_mode(args, dev, local):
    from lush_nerf_b200 import _lib, ops
    peaks = measured_peaks()
    R = 65536
    results = []
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()

    def timed(fns, iters):
        """fns: one closure per rotating input set; device time per call (CUDA events on the launching stream)."""
        for f in fns:
            f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.cuda.CUDAGraph()          # the calls are short: one graph of `iters` launches, replay timed
        with torch.cuda.graph(g):
            for i in range(iters):
                fns[i % len(fns)]()
        g.replay()
        torch.cuda.synchronize()
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e3          # us

    for S in (32, 64, 128, 256):
        gen = torch.Generator(device=dev).manual_seed(S)
        per_set = R * S * 4 * 8                          # raw + z + noise + outputs of one compositing call, roughly
        nset = max(2, int(np.ceil(2.5 * L2_BYTES / per_set)))
        sets = []
        for _ in range(nset):
            raw = torch.randn(R, S, 4, device=dev, generator=gen)
            raw[..., 3] = torch.relu(3 * torch.randn(R, S, device=dev, generator=gen)) * 5      # peaky densities
            z = (torch.linspace(0, 1, S, device=dev).expand(R, S)
                 + torch.rand(R, S, device=dev, generator=gen) * (0.5 / S)).contiguous()
            rd = torch.randn(R, 3, device=dev, generator=gen)
            nz = torch.randn(R, S - 1, device=dev, generator=gen)
            sets.append(dict(raw=raw, z=z, rd=rd, nz=nz))
        iters = max(2 * nset, 12)

        def rec(name, us, nbytes, extra=None):
            gbs = nbytes / us / 1e3
            d = {"kernel": name, "R": R, "S": S, "N": S, "us": us, "algorithmic_bytes": nbytes, "GBs": gbs,
                 "frac_of_hbm": gbs / peaks["hbm_gbs"]}
            d.update(extra or {})
            results.append(d)

        with torch.no_grad():
            us = timed([lambda s=s: ops.composite(s["raw"], s["z"], s["rd"], s["nz"]) for s in sets], iters)
        rec("composite_fwd", us, R * (32 * S + 24))
        # backward through the C ABI entry itself (the call autograd's backward makes), incoming gradient of rgb_map only
        from ctypes import c_void_p
        outs = [(torch.randn(R, 3, device=dev, generator=gen), torch.empty(R, S, 4, device=dev)) for _ in sets]

        def bwd(s, o):
            _lib.check(_lib.lib().lush_composite_bwd(
                _lib.ptr(s["raw"]), _lib.ptr(s["z"]), c_void_p(s["rd"].data_ptr()), 3, _lib.ptr(s["nz"]), -1.0, 0,
                _lib.ptr(o[0]), None, None, None, None, _lib.ptr(o[1]), None, R, S, _lib.stream()), "lush_composite_bwd")
        us = timed([lambda s=s, o=o: bwd(s, o) for s, o in zip(sets, outs)], iters)
        rec("composite_bwd", us, R * (40 * S + 40))
        with torch.no_grad():
            ws = [ops.composite(s["raw"], s["z"], s["rd"], s["nz"])[3] for s in sets]
            us_r = [torch.rand(R, S, device=dev, generator=gen) for _ in sets]
            us = timed([lambda s=s, w=w, u=u: ops.sample_merge(s["z"], w, u, strict=False)
                        for s, w, u in zip(sets, ws, us_r)], iters)
            rec("sample_merge(rand, fused cdf)", us, R * (4 * S + 4 * S + 4 * S + 4 * 2 * S + 4 * S),
                {"bytes": "z 4S + weights 4S + u 4N in, z_fine 4(S+N) + z_samples 4N out"})
            bins = [(0.5 * (s["z"][:, 1:] + s["z"][:, :-1])).contiguous() for s in sets]
            wm = [w[:, 1:-1].contiguous() for w in ws]
            ud = ops.linspace01(S, dev)
            us = timed([lambda b=b, w=w: ops.sample_pdf_cuda(b, w, ud) for b, w in zip(bins, wm)], iters)
            rec("sample_pdf(det, fused cdf)", us, R * (8 * S + 4 * S - 12))
            us = timed([lambda b=b, w=w, u=u: ops.sample_pdf_cuda(b, w, u) for b, w, u in zip(bins, wm, us_r)], iters)
            rec("sample_pdf(rand, fused cdf)", us, R * (8 * S + 8 * S - 12))
            cdfs = [ops.reference_cdf(w) for w in wm]
            us = timed([lambda b=b, w=w, c=c: ops.sample_pdf_cuda(b, w, ud, cdf=c) for b, w, c in zip(bins, wm, cdfs)], iters)
            rec("sample_pdf(det, strict: search kernel only, cdf given)", us, R * (8 * S + 4 * S - 12))
        del sets, outs, ws, us_r, bins, wm, cdfs
        torch.cuda.empty_cache()
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    top = next(r for r in results if r["kernel"] == "composite_fwd" and r["S"] == 128)
    line = {"metric": "compositing fwd GB/s (65536 rays, S=128)", "value": top["GBs"], "unit": "GB/s", "n_gpus": 1,
            "steps": 1, "warmup": 1, "ms_per_step": top["us"] / 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "stress sweep (BASELINE configs[4]): 65 536-ray batches x S = N in {32,64,128,256}, "
                                   "compositing fwd/bwd, sample_pdf, sample_merge; LLFF-like peaky densities",
                       "l2": "inputs rotate through > 2.5x L2 of distinct buffers inside one graph replay"},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": top["GBs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": top["frac_of_hbm"], "traffic": None,
                         "kernel": "composite_fwd_kernel, algorithmic bytes 32 S + 24 per ray", "peak_source": peaks["source"]},
            "kernels": results}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "stress"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--mlp", default=None, choices=[None, "simt", "tc"], help="override LUSH_MLP_IMPL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", default="single", choices=["single", "split", "none"],
                    help="CUDA-graph capture of the step: one graph (NCCL inside for N>1), two graphs around eager "
                         "collectives, or eager launches")
    ap.add_argument("--no-graph", action="store_true", help="same as --graph none")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--render-frames", type=int, default=10)
    ap.add_argument("--table-batches", action="store_true",
                    help="feed the reference's batch form (rays sliced from a host ray table) instead of index batches")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.no_graph:
        args.graph = "none"
    if args.mlp:
        os.environ["LUSH_MLP_IMPL"] = args.mlp
    args.warmup = max(args.warmup, 3)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.mode == "stress":
        if rank == 0:
            stress_mode(args, dev, local)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    os.environ.setdefault("LUSH_VERBOSE", "0")

    from lush_nerf_b200 import _lib, ops
    from lush_nerf_b200.synthetic import build_model, poster_args
    from lush_nerf_b200.train import Trainer, shard_rows

    torch.manual_seed(0)
    np.random.seed(0)
    pargs = poster_args()
    model = build_model(pargs).to(dev)
    if world > 1:  # identical replicas: broadcast rank 0's initial parameters once
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    torch.manual_seed(1000 + rank)      # the ranks draw different jitter / sigma-noise streams
    render_kwargs = dict(perturb=1.0, N_importance=N_IMPORTANCE, N_samples=N_SAMPLES, use_viewdirs=True,
                         white_bkgd=False, raw_noise_std=1.0, inference=False, near=0.0, far=1.0)
    scene = Scene(0)
    trainer = Trainer(model, H, W, scene.K, lrate=5e-4, render_kwargs=render_kwargs)
    trainer.set_scene(scene.poses, scene.images, scene.masks)
    strong = args.scaling == "strong" and world > 1
    n_local = N_RAND // world if strong else N_RAND
    n_batches = args.warmup + args.steps
    host = []
    for i in range(n_batches):
        if strong:       # ONE global batch per step (same generator on every rank), each rank keeps its rows
            b = scene.index_batch(N_RAND, torch.Generator().manual_seed(10 + i))
            lo, hi = shard_rows(N_RAND, rank, world)
            b = {k: v[lo:hi].contiguous() for k, v in b.items()}
        else:
            b = scene.index_batch(N_RAND, torch.Generator().manual_seed(10 + i + 1000 * rank))
        if args.table_batches:
            b = scene.expand(b)
        host.append({k: v.contiguous().pin_memory() for k, v in b.items()})
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- MLP timing hook (roofline): CUDA events around every lush_mlp_fwd / lush_mlp_bwd call
    mlp_events = []
    ops.MLP_EVENT_SINK = None

    # ---- warm-up (eager); the clock sampler starts here (nvidia-smi needs ~0.3 s before its first sample) and runs
    # through the timed regions: every sample is taken under load
    sampler = ClockSampler(local, period_ms=50)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        trainer.step(resident[i])
    barrier()

    def timed_pass(with_events):
        """K steps over batches already resident in HBM; returns device ms (CUDA events on the launching stream)."""
        ops.MLP_EVENT_SINK = mlp_events if with_events else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            trainer.step(resident[args.warmup + i])
        e1.record()
        barrier()
        ops.MLP_EVENT_SINK = None
        return e0.elapsed_time(e1)

    # ---- timed region 1a: eager launches with CUDA events around every MLP call (roofline numerators, launch count)
    launches0 = _lib.launch_count()
    ms_eager = timed_pass(True)
    launches = _lib.launch_count() - launches0
    def _ms(labels):
        return sum(a.elapsed_time(b) for lab, a, b in mlp_events if lab in labels)
    mlp_ms = _ms(("fwd", "bwd", "fwd_simt", "bwd_simt"))
    fwd_ms, bwd_ms = _ms(("fwd",)), _ms(("bwd",))
    simt_ms = _ms(("fwd_simt", "bwd_simt"))
    mlp_calls = len(mlp_events)

    # ---- timed region 1b: the same step captured in ONE CUDA graph (Trainer.enable_cuda_graph) -- the product path
    graphed = False
    if args.graph != "none":
        try:
            graphed = bool(trainer.enable_cuda_graph(resident[0], mode=args.graph))
            for i in range(args.warmup):
                trainer.step(resident[i])
        except Exception as e:   # report, never hide: the eager path is the same kernels, only launched one by one
            print(f"[bench] CUDA graph capture failed, running eagerly: {type(e).__name__}: {e}", file=sys.stderr)
            graphed = False
            trainer._graph = trainer._graph_tail = None
    if world > 1:   # all ranks take the same path (a rank that failed to capture would still match collectives, but
        ok = torch.tensor([1 if graphed else 0], device=dev)        # the line should describe ONE configuration)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if graphed and not bool(ok.item()):
            graphed = False
            trainer._graph = trainer._graph_tail = None
    ms_total = timed_pass(False) if graphed else ms_eager
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end from pinned host buffers, loss read back every step
    barrier()
    t0 = time.perf_counter()
    # Every step: H2D of the step's batch from pinned memory (straight into the captured step's static input buffers
    # when the step is a graph) and a D2H read of its loss.  The read is pipelined like a training loop's logging: the
    # loss of step i is copied into pinned memory asynchronously and consumed while step i+1 is already queued; the
    # last one is read synchronously inside the timed region.
    loss_pin = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    losses_host = []
    for i in range(args.steps):
        hb = host[args.warmup + i]
        b = hb if graphed else {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
        loss = trainer.step(b)
        loss_pin[i & 1].copy_(loss.detach(), non_blocking=True)
        loss_ev[i & 1].record()
        if i > 0:
            loss_ev[(i - 1) & 1].synchronize()
            losses_host.append(float(loss_pin[(i - 1) & 1]))
    loss_ev[(args.steps - 1) & 1].synchronize()
    losses_host.append(float(loss_pin[(args.steps - 1) & 1]))
    loss_host = losses_host[-1]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if not os.environ.get("LUSH_BENCH_ALLOW_NONFINITE"):      # (timing experiments with deliberately broken kernels)
        assert len(losses_host) == args.steps and all(np.isfinite(losses_host)), "e2e: a step's loss did not come back finite"
    guard_bits = trainer.check_numerics()

    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    replicas_in_sync = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # identical Adam on identical all-reduced gradients: the replicas must still hold the same parameters
        flat = torch.cat([p.detach().reshape(-1).double() for p in model.parameters()])
        chk = torch.stack([flat.sum(), (flat * flat).sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        replicas_in_sync = bool(torch.equal(lo, hi))
    ms_total, e2e_ms = t.tolist()
    rays_total = n_local * K_RAYS * world * args.steps
    value = rays_total / (ms_total * 1e-3)
    e2e_val = rays_total / (e2e_ms * 1e-3)

    # ---- render leg (BASELINE configs[3]): --render_only style full frames, chunk 32768, 64+64, det sampling
    render = None
    if not args.no_render:
        try:   # a failure of this secondary leg must not cost the headline line (all ranks fail or succeed together:
            render = render_leg(args, model, scene.K, ops, mlp_events, dev, local, rank, world, barrier)   # same code)
        except Exception as e:
            render = {"error": f"{type(e).__name__}: {e}"}
        model.train()

    if rank == 0:
        peaks = measured_peaks()
        flops_step = 3 * FLOP_FWD_PER_SAMPLE * n_local * K_RAYS * SAMPLES_PER_RAY
        tc = model.mlp_coarse.impl() == 1
        scene_ms = (fwd_ms + bwd_ms) if tc else mlp_ms
        achieved = (flops_step * args.steps / (scene_ms * 1e-3)) / 1e12 if scene_ms > 0 else None
        impl_name = {0: "fp32 SIMT (FFMA)", 1: "tcgen05 fp16 operands / fp32 accumulate"}[model.mlp_coarse.impl()]
        # per launch class (CUDA events around the C-ABI calls on the launching stream, timed region only)
        per_kernel = None
        if tc and fwd_ms > 0 and bwd_ms > 0:
            f = flops_step / 3 * args.steps
            per_kernel = {
                "mlp_fwd (pack + mlp_tc_fwd_kernel, stash written)": {
                    "ms_per_step": fwd_ms / args.steps, "tflops": f / (fwd_ms * 1e-3) / 1e12,
                    "frac_of_tensor_peak": f / (fwd_ms * 1e-3) / 1e12 / peaks["tflops"]},
                "mlp_bwd (dgrad chain incl. head gradients + wgrad GEMM + ray reductions)": {
                    "ms_per_step": bwd_ms / args.steps, "tflops": 2 * f / (bwd_ms * 1e-3) / 1e12,
                    "frac_of_tensor_peak": 2 * f / (bwd_ms * 1e-3) / 1e12 / peaks["tflops"]},
                "noise MLP (fp32 SIMT, 4x128, 1024 rows)": {"ms_per_step": simt_ms / args.steps},
            }
        traffic = load_traffic()
        graph_desc = (trainer.graph_mode if graphed else "none")
        line = {
            "metric": "train rays/s (coarse+fine fwd+bwd, 64+64 samples)", "value": value, "unit": "rays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "fp32" if model.mlp_coarse.impl() == 0 else "fp16 operands, fp32 accumulate",
            "data": "synthetic", "config": workload_config(world, "strong" if strong else "weak", graph_desc),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "cuda_graph": graphed, "ms_per_step_eager": ms_eager / args.steps, "replicas_in_sync": replicas_in_sync,
            "pixels_per_s": value / K_RAYS, "guard_bits": guard_bits,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["tflops"]) if achieved else None,
                         "frac_of_burst_peak": (achieved / peaks["tflops_burst"]) if achieved else None,
                         "traffic": traffic.get("bytes_per_step") if traffic else None,
                         "traffic_source": traffic.get("source") if traffic else None,
                         "per_kernel": per_kernel,
                         "kernel": f"encoder+MLP fwd+bwd ({impl_name}), all launches of the step; algorithmic "
                                   f"{flops_step / 1e12:.3f} TFLOP/step (3 x 1 186 816 FLOP x "
                                   f"{n_local * K_RAYS * SAMPLES_PER_RAY} samples)",
                         "mlp_ms_per_step": scene_ms / args.steps, "mlp_calls_per_step": mlp_calls / args.steps,
                         "peak_source": peaks["source"]},
            "final_loss": loss_host,
            "render": render,
        }
        line["cpu_baseline"] = None
        if not args.no_cpu_baseline and world == 1:
            try:
                val, ms, cores, sample = cpu_reference_run(steps=2, warmup=1, budget_s=25.0)
                line["cpu_baseline"] = {"value": val, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample}
            except Exception as e:
                line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # A CUDA graph that holds captured NCCL kernels must be gone before the communicator is torn down (observed on
        # 2 x B200: destroy_process_group() never returned with the graph alive); and nothing after this point is worth
        # a hang: the line is printed, leave without the interpreter's teardown.
        trainer._graph = trainer._graph_tail = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
