#!/usr/bin/env python
"""bench.py -- train rays/s of the LuSh-NeRF per-ray rendering hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode train|render]

One "step" = one pass of the hot path over one batch of synthetic LLFF-shaped rays at the
poster_lushnerf training configuration (BASELINE.json configs[1]): NeRFAll.forward in rigid-blur-kernel
mode (N_rand=1024 pixels x K=5 warped rays = 5120 rays, 64 coarse + 64 fine samples, perturb=1,
raw_noise_std=1) + loss + backward + gradient all-reduce + Adam.  For N>1 every rank renders its own
N_rand pixels (weak scaling) and the only collective is one flat 5.2 MB gradient all-reduce.

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with the batch already resident in HBM;
`e2e` = the same step driven from pinned host buffers (H2D of the batch + D2H of the loss inside the timed
region); `roofline` = algorithmic MLP FLOPs / measured MLP device time against the measured bf16 peak;
`cpu_baseline` = the CPU oracle port on this box's host cores on a bounded sample.
`--impl reference` times the reference algorithm's CPU implementation (oracle port; the reference tree
itself cannot travel to the GPU box) on the host cores for the same metric.
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


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "tflops": d.get("bf16_tflops_sustained", 1400.0),
                "source": "MEASURED_PEAKS.json (bf16 sustained: kernel timed inside a long step)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def load_traffic():
    """dram bytes (read + write) of the scene-MLP kernels of ONE training step, from the committed ncu
    `--set full` capture of this very command (profiles/r01_ncu_mlp_step.json, written by tools/ncu_step.py)."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_mlp_step.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return {"bytes_per_step": d["dram_bytes_per_step"], "source": "profiles/r01_ncu_mlp_step.json"}
    except Exception:
        return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=open(self.path, "w"),
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
# CPU arm: the reference algorithm (oracle port) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, n_pixels=32):
    """fwd+bwd of the poster training step on the CPU oracle; returns (rays/s, ms/step, cores, sample)."""
    from oracle import lush_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nets = O.det_nets(O.POSTER_CFG, tag=1, sigma_gain=30.0)
    rbk = O.det_rbk_params(tag=7)
    pn = {k: {n: t.clone().requires_grad_(True) for n, t in d.items()} for k, d in nets.items()}
    pr = {n: t.clone().requires_grad_(True) for n, t in rbk.items()}
    from lush_nerf_b200.synthetic import SyntheticLLFF
    data = SyntheticLLFF(H, W, FOCAL, max_views=1)
    R = n_pixels * K_RAYS
    times = []
    for it in range(warmup + steps):
        b = data.batch(it, n_pixels)
        t0 = time.perf_counter()
        out, _ = O.nerfall_train_forward(
            pn, pr, O.POSTER_CFG, b["rays"].contiguous(), b["images_idx"], H, W, FOCAL, N_samples=N_SAMPLES,
            N_importance=N_IMPORTANCE, allkernel=True, kernel_pixel=b["fq_mask"], t_rand=torch.rand(R, N_SAMPLES),
            noise_c=torch.randn(R, N_SAMPLES - 1), noise_f=torch.randn(R, N_SAMPLES + N_IMPORTANCE - 1),
            u=torch.rand(R, N_IMPORTANCE))
        tgt = b["rgbsf"]
        loss = 0.5 * ((out[0] - tgt) ** 2).mean() + 0.5 * (out[0] - tgt).abs().mean() \
            + 0.5 * ((out[1] - tgt) ** 2).mean() + 0.5 * (out[1] - tgt).abs().mean()
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    sample = (f"{n_pixels} pixels x {K_RAYS} rays = {R} rays/step, {N_SAMPLES}+{N_IMPORTANCE} samples, "
              f"NeRFAll.forward RBK mode fwd+bwd, {steps} timed steps, torch CPU fp32")
    return R * len(times) / total, 1e3 * total / len(times), cores, sample


def render_leg(args, model, trainer, ops, mlp_events, dev, local):
    """ms/frame of NeRFAll.forward in eval mode over a spiral path (rays generated from the pose on the device)."""
    from lush_nerf_b200.synthetic import spiral_poses
    poses = torch.from_numpy(spiral_poses(args.render_frames + 2, radius_scale=2.0)).to(dev)
    rk = dict(perturb=False, N_importance=N_IMPORTANCE, N_samples=N_SAMPLES, use_viewdirs=True, white_bkgd=False,
              raw_noise_std=0.0, inference=True, near=0.0, far=1.0)
    model.eval()
    with torch.no_grad():
        model(H, W, trainer.K, chunk=1024 * 32, poses=poses[:2], render_kwargs=dict(rk))   # warm-up frames
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rsampler = ClockSampler(local)
        rsampler.start()
        mlp_events.clear()
        ops.MLP_EVENT_SINK = mlp_events
        try:
            r0.record()
            rgbs, _, _ = model(H, W, trainer.K, chunk=1024 * 32, poses=poses[2:], render_kwargs=dict(rk))
            r1.record()
            torch.cuda.synchronize()
        finally:
            ops.MLP_EVENT_SINK = None
    rclocks = rsampler.stop()
    ms_frame = r0.elapsed_time(r1) / args.render_frames
    mlp_ms_frame = sum(a.elapsed_time(b) for lab, a, b in mlp_events if lab == "fwd") / args.render_frames
    fl = FLOP_FWD_PER_SAMPLE * H * W * SAMPLES_PER_RAY
    return {"ms_per_frame": ms_frame, "frames": args.render_frames, "rays_per_s": H * W / (ms_frame * 1e-3),
            "image": f"{H}x{W}", "chunk": 32768, "mlp_ms_per_frame": mlp_ms_frame,
            "mlp_tflops": fl / (mlp_ms_frame * 1e-3) / 1e12 if mlp_ms_frame > 0 else None,
            "finite": bool(torch.isfinite(rgbs).all().item()), "clocks": rclocks}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    val, ms, cores, sample = cpu_reference_run(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "train rays/s (coarse+fine fwd+bwd, 64+64 samples)", "value": val,
            "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(1),
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "poster_lushnerf training step (BASELINE configs[1]): NeRFAll.forward RBK mode, "
                        "N_rand=1024 px x K=5 rays = 5120 rays/GPU, N_samples=64, N_importance=64, perturb=1, "
                        "raw_noise_std=1, scene+noise MLPs, loss, backward, Adam",
            "rays_per_step_per_gpu": N_RAND * K_RAYS, "image": f"{H}x{W} f={FOCAL}", "views": 20,
            "parallelism": f"rays sharded x{n_gpus}, one flat grad all-reduce/step" if n_gpus > 1 else "single GPU",
            "l2": "per-step working set (GBs of activation stash) exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mlp", default=None, choices=[None, "simt", "tc"], help="override LUSH_MLP_IMPL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the training step in a CUDA graph")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--render-frames", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.mlp:
        os.environ["LUSH_MLP_IMPL"] = args.mlp
    args.warmup = max(args.warmup, 3)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    os.environ.setdefault("LUSH_VERBOSE", "0")

    from lush_nerf_b200 import _lib, ops
    from lush_nerf_b200.synthetic import SyntheticLLFF, build_model, poster_args
    from lush_nerf_b200.train import Trainer

    torch.manual_seed(0)
    np.random.seed(0)
    pargs = poster_args()
    model = build_model(pargs).to(dev)
    if world > 1:  # identical replicas: broadcast rank 0's initial parameters once
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    render_kwargs = dict(perturb=1.0, N_importance=N_IMPORTANCE, N_samples=N_SAMPLES, use_viewdirs=True,
                         white_bkgd=False, raw_noise_std=1.0, inference=False, near=0.0, far=1.0)
    trainer = Trainer(model, H, W, np.array([[FOCAL, 0, 0.5 * W], [0, FOCAL, 0.5 * H], [0, 0, 1]]), lrate=5e-4,
                      render_kwargs=render_kwargs)
    data = SyntheticLLFF(H, W, FOCAL, seed=rank, max_views=2)
    n_batches = args.warmup + args.steps
    host = [{k: v.contiguous().pin_memory() for k, v in data.batch(i, N_RAND).items()} for i in range(n_batches)]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- MLP timing hook (roofline): CUDA events around every lush_mlp_fwd / lush_mlp_bwd call
    mlp_events = []
    ops.MLP_EVENT_SINK = None

    # ---- warm-up (eager)
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
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
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
    # N > 1: two graphs around an eager NCCL all-reduce (capturing the collective itself hung on 2 x B200 in round 1)
    if not args.no_graph:
        try:
            graphed = bool(trainer.enable_cuda_graph(resident[0]))
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
    assert len(losses_host) == args.steps and all(np.isfinite(losses_host)), "e2e: a step's loss did not come back finite"

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
    rays_total = N_RAND * K_RAYS * world * args.steps
    value = rays_total / (ms_total * 1e-3)
    e2e_val = rays_total / (e2e_ms * 1e-3)

    # ---- render leg (BASELINE configs[3]): --render_only style full frames, chunk 32768, 64+64, det sampling
    render = None
    if world == 1 and not args.no_render:
        try:   # a failure of this secondary leg must not cost the headline line
            render = render_leg(args, model, trainer, ops, mlp_events, dev, local)
        except Exception as e:
            render = {"error": f"{type(e).__name__}: {e}"}
        model.train()

    if rank == 0:
        peaks = measured_peaks()
        flops_step = 3 * FLOP_FWD_PER_SAMPLE * N_RAND * K_RAYS * SAMPLES_PER_RAY
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
                "mlp_bwd (dgrad chain + wgrad GEMM + head/ray reductions)": {
                    "ms_per_step": bwd_ms / args.steps, "tflops": 2 * f / (bwd_ms * 1e-3) / 1e12,
                    "frac_of_tensor_peak": 2 * f / (bwd_ms * 1e-3) / 1e12 / peaks["tflops"]},
                "noise MLP (fp32 SIMT, 4x128, 1024 rows)": {"ms_per_step": simt_ms / args.steps},
            }
        traffic = load_traffic()
        line = {
            "metric": "train rays/s (coarse+fine fwd+bwd, 64+64 samples)", "value": value, "unit": "rays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32" if model.mlp_coarse.impl() == 0 else "fp16 operands, fp32 accumulate",
            "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "cuda_graph": graphed, "ms_per_step_eager": ms_eager / args.steps, "replicas_in_sync": replicas_in_sync,
            "pixels_per_s": value / K_RAYS,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["tflops"]) if achieved else None,
                         "traffic": traffic.get("bytes_per_step") if traffic else None,
                         "traffic_source": traffic.get("source") if traffic else None,
                         "per_kernel": per_kernel,
                         "kernel": f"encoder+MLP fwd+bwd ({impl_name}), all launches of the step; algorithmic "
                                   f"{flops_step / 1e12:.3f} TFLOP/step (3 x 1 186 816 FLOP x 983 040 samples)",
                         "mlp_ms_per_step": scene_ms / args.steps, "mlp_calls_per_step": mlp_calls / args.steps,
                         "peak_source": peaks["source"]},
            "final_loss": loss_host,
            "render": render,
        }
        line["cpu_baseline"] = None
        if not args.no_cpu_baseline and world == 1:
            try:
                val, ms, cores, sample = cpu_reference_run(steps=2, warmup=1)
                line["cpu_baseline"] = {"value": val, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample}
            except Exception as e:
                line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
