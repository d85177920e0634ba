#!/usr/bin/env python
"""Headline benchmark of the NRHints ray-march hot path (BASELINE.json config #2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rays R] [--mlp auto|fp32|tcgen05]

A step = one NeuSHintRenderer.forward over 4096 rays x 128 samples (64 coarse + 64 importance, one shadow
ray of 64 + 64 samples per primary ray, both hints, white background, inference mode) of the synthetic
800x800 workload (nrhints_b200/workload.py), random-init (geometric) weights at seed 3407.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.

--impl reference times the reference algorithm on the host cores: the reference is pure Python/PyTorch and
cannot travel to the GPU box, so the timed code is the oracle port (oracle/nrh_oracle.py, pinned to the
reference by the golden fixtures), on all host threads, on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

FLOP_PER_RAY = 766.6e6          # algorithmic, de-duplicated forward FLOPs per ray
BYTES_PER_RAY = 7240            # 52 B in + 7188 B out
# dominant kernel = the fine-pass SDF kernel (forward with feature head + reverse sweep) over R*128 points
FLOP_PER_FINE_POINT = 1049088.0 + 918016.0


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


_BEST_THREADS = None


def _pick_threads():
    """The reference runs PyTorch with min(8, cpus/gpus) threads (trainer/launcher.py:43); on a many-core host more
    threads help only up to a point and over-subscription hurts, so calibrate once on a small sample and keep the best."""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    cores = os.cpu_count() or 1
    best, best_t = None, None
    for th in sorted({min(cores, t) for t in (8, 16, 32, 64, cores)}):
        _, med = cpu_reference_leg(32, 0, threads=th)
        if best_t is None or med < best_t:
            best, best_t = th, med
    _BEST_THREADS = best
    return best


def cpu_reference_leg(n_rays: int, repeats: int, threads: int = None):
    """The reference algorithm (oracle port) on the host cores on `n_rays` rays of the workload."""
    from oracle import nrh_oracle as orc
    import nrhints_b200 as nb
    from nrhints_b200.workload import synthetic_rays
    cores = threads if threads is not None else _pick_threads()
    torch.set_num_threads(cores)
    cfg = nb.NeuSModelConfig()
    torch.manual_seed(3407)
    state = {k: v.detach().clone() for k, v in nb.NeuSHintRenderer(cfg).state_dict().items()}
    ocfg = orc.OracleConfig.from_model_config(cfg)
    rays = synthetic_rays(n_rays, seed=3407)
    bg = torch.ones(1, 3)
    times = []
    for i in range(repeats + 1):                   # first pass = warm-up
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.render_forward(state, ocfg, rays["origins"], rays["directions"], rays["pl_positions"], rays["nears"],
                               rays["fars"], background_rgb=bg)
        times.append(time.perf_counter() - t0)
    times = sorted(times[1:]) if repeats > 0 else times
    med = times[len(times) // 2]
    return {"value": n_rays / med, "unit": "rays/s", "cores": cores, "kind": "port", "host_cores": os.cpu_count(),
            "sample": f"{n_rays} rays x 128 samples of the same workload, 1 warm-up + median of {max(repeats, 1)} passes, "
                      f"torch {torch.get_num_threads()} threads = best of {{8,16,32,64,all}} on this host "
                      f"(oracle port of the reference PyTorch path)"}, med


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_rays = 256
    per_step = []
    base, _ = cpu_reference_leg(n_rays, 0)           # warm-up pass
    from oracle import nrh_oracle as orc  # noqa: F401
    for _ in range(max(args.warmup - 1, 0)):
        cpu_reference_leg(n_rays, 0)
    vals = []
    for _ in range(args.steps):
        b, med = cpu_reference_leg(n_rays, 0)
        vals.append(med)
    ms = 1e3 * sum(vals) / len(vals)
    value = n_rays / (sum(vals) / len(vals))
    base.update(value=value)
    line = {"impl": "reference", "metric": "rays/sec (4096 rays x 128 samples forward render)", "value": value, "unit": "rays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "NRHints forward render, 64+64 samples, shadow 64+64, both hints, 800x800 synthetic scene, "
                                   f"bounded sample of {n_rays} rays per step on host CPU", "rays_per_step": n_rays},
            "cpu_baseline": base, "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner to fd 1 from every
    rank), so fd 1 is pointed at stderr for the whole run and the result line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--mlp", default="auto", choices=["auto", "fp32", "tcgen05"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step measurement (config #3)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import nrhints_b200 as nb
    from nrhints_b200 import _lib
    from nrhints_b200.workload import synthetic_rays

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    R = args.rays

    cfg = nb.NeuSModelConfig()
    torch.manual_seed(3407)
    model = nb.NeuSHintRenderer(cfg, mlp_impl=args.mlp).to(dev)
    host_rays = nb.RayBundle(**synthetic_rays(R, seed=3407 + rank)).pin_memory()      # each rank renders its own rays
    dev_rays = host_rays.to(dev)
    bg = torch.ones(1, 3, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                       # > 126 MB L2

    torch.set_grad_enabled(False)          # config #2 is the inference render (the reference evaluates under @torch.no_grad(),
                                           # pipelines/base_pipeline.py:93); with grad enabled the module would add its autograd backend

    def step():
        return model(dev_rays, is_training=False, background_rgb=bg)

    for _ in range(W):
        step()
    torch.cuda.synchronize()
    launches_per_step = model.last_launch_count + 5          # + the torch ops of the wrapper (s_val, inv_s)
    lib = _lib.load()
    import ctypes as C
    ccfg = model._c_config()
    engine = "tcgen05-fp16x3" if (args.mlp == "tcgen05" or (args.mlp == "auto" and _engine_is_tc(lib, ccfg))) else "fp32-simt"

    # ---- timed region: K steps, CUDA events per step on the launching stream, L2 flushed between steps ----------
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall = time.perf_counter() - t_wall0
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / K
    value = world * R / (ms_per_step * 1e-3)

    # ---- e2e: host (pinned) rays -> device -> forward -> full RenderOutput back on the host -------------------
    e2e_steps = max(3, min(K, 5))
    out_host = None
    for _ in range(2):
        out_host = model.render_to_host(host_rays, background_rgb=bg)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out_host = model.render_to_host(host_rays, background_rgb=bg)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d = sum(v.numel() * 4 for v in (host_rays.origins, host_rays.directions, host_rays.pl_positions, host_rays.nears, host_rays.fars))
    d2h = sum(v.numel() * v.element_size() for k, v in out_host.as_dict().items() if v is not None and k != "relax_inside_sphere")

    # ---- roofline of the dominant kernel: fine-pass SDF kernel (forward + feature head + reverse sweep) --------
    roof = None
    if rank == 0:
        peaks = measured_peaks()
        pts = (torch.rand(R * 128, 3, device=dev) - 0.5) * 2.0
        for _ in range(2):
            model.sdf_query(pts, want_grad=True, want_feat=True)
        torch.cuda.synchronize()
        ks = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            lib_rc = model.sdf_query(pts, want_grad=True, want_feat=True)
            b.record()
            torch.cuda.synchronize()
            ks.append(a.elapsed_time(b))
            del lib_rc
        k_ms = sum(ks) / len(ks)          # includes three tiny torch.empty allocations (cached allocator, no kernel)
        achieved = FLOP_PER_FINE_POINT * R * 128 / (k_ms * 1e-3) / 1e12
        # the kernel is timed ALONE (one launch between L2 flushes), so its denominator is the burst cuBLAS figure; the whole step
        # (a long back-to-back run) is quoted against the sustained one
        peak, peak_sus = peaks["bf16_tflops"], peaks["bf16_tflops_sustained"]
        roof = {"bound": "tensor", "kernel": "sdf_mlp (fine pass: forward + feature head + reverse sweep)", "engine": engine,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_kind": f"cuBLAS bf16 dense, burst (kernel timed alone), {peaks['source']} (MEASURED_PEAKS.json); the fp16x3 split issues 3 "
                             "MMAs per logical product (its own ceiling is 1/3 of the fp16 MMA rate: 469 TFLOP/s logical), the fp32-simt "
                             "engine runs on FFMA (75 TFLOP/s nominal)",
                "frac_of_sustained_peak": achieved / peak_sus,
                "kernel_ms": k_ms, "flop_per_launch": FLOP_PER_FINE_POINT * R * 128, "traffic": (_ncu_traffic(engine) or {}).get("bytes_per_launch"), "traffic_unit": "B per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                "traffic_detail": _ncu_traffic(engine),
                "whole_step": {"achieved_tflops": FLOP_PER_RAY * R / (ms_per_step * 1e-3) / 1e12,
                               "frac_of_tensor_peak": FLOP_PER_RAY * R / (ms_per_step * 1e-3) / 1e12 / peak_sus,
                               "hbm_algorithmic_gbs": BYTES_PER_RAY * R / (ms_per_step * 1e-3) / 1e9,
                               "frac_of_hbm_peak": BYTES_PER_RAY * R / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"]}}

    # ---- BASELINE.json config #3: one training step (forward + loss + backward + Adam) on the same 4096 rays ---------
    train = None
    if not args.no_train:
        torch.set_grad_enabled(True)
        from types import SimpleNamespace
        from nrhints_b200.grad_sync import allreduce_flat
        from nrhints_b200.workload import synthetic_pixel_bundle
        # the reference's train_iter (trainer/trainer.py:269-283) on the composed pipeline: pixel bundle -> ray generation ->
        # render -> loss dict -> backward -> Adam; the renderer is the one measured above (same weights)
        pb, cam = synthetic_pixel_bundle(R, seed=3407 + rank)
        pipe = nb.NRHintPipeline(cfg, nb.RayGeneratorConfig(), nb.CameraModel(**cam), 64, mlp_impl=args.mlp)
        pipe.renderer = model
        pipe = pipe.to(dev)
        pixels = SimpleNamespace(**{k: v.to(dev) for k, v in vars(pb).items()})
        opt = pipe.make_optimizer()                              # parameters / gradients / moments re-homed into flat buffers

        def train_step():
            opt.zero_grad()                                        # one memset
            res = pipe(pixels, global_step=60000)
            loss = pipe.get_train_loss_dict(res, pixels)["loss"]  # pipelines/base_pipeline.py:57-62, 2 launches
            loss.backward()
            # the one collective of the training loop: the flat gradient buffer itself (no packing); the mean over ranks is
            # folded into the Adam launch
            scale = allreduce_flat(opt.flat_grads()) if dist is not None else 1.0
            opt.step(grad_scale=scale)
        for _ in range(2):
            train_step()
        torch.cuda.synchronize()
        tsteps = max(3, min(K, 5))
        tev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(tsteps)]
        if dist is not None:
            dist.barrier()
        for a, b in tev:
            flush.zero_()
            a.record(); train_step(); b.record()
        torch.cuda.synchronize()
        tt = torch.tensor([sum(a.elapsed_time(b) for a, b in tev) / tsteps], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = float(tt.item())
        train = {"value": world * R / (t_ms * 1e-3), "unit": "rays/s", "ms_per_step": t_ms, "steps": tsteps,
                 "what": "BASELINE config #3: ray generation + forward + L1/eikonal loss + backward + Adam on 4096 rays/GPU (the reference's train_iter), is_training=True (jitter, "
                         "global_step 60000); fused CUDA SDF forward-with-tape / backward (tcgen05), fused loss (2 launches) and flat-buffer Adam (1 launch); "
                         "CUDA compositing node (forward + hand-derived backward), reflectance MLP on fp16 library GEMMs"
                         + ("; flat-buffer gradient all-reduce" if dist is not None else "")}
        torch.set_grad_enabled(False)
        del opt

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference_leg(256, 1)

    if rank == 0:
        line = {
            "metric": "rays/sec (4096 rays x 128 samples forward render)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if engine == "fp32-simt" else "f32 (fp16 hi/lo split operands, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "NRHints NeuSHintRenderer.forward, BASELINE config #2: 4096 rays x (64+64) samples, shadow ray 64+64, "
                                   "shadow+specular hints, white bg, 800x800 synthetic scene, geometric-init weights seed 3407",
                       "rays_per_gpu": R, "samples_per_ray": 128, "mlp_engine": engine, "parallelism": f"rays sharded x{world}, no data-path collective",
                       "l2": "256 MiB buffer rewritten between timed steps (L2 flush); per-step working set ~0.9 GB > 126 MB L2"},
            "e2e": {"value": world * R / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "pinned host RayBundle -> device, forward, full RenderOutput -> pinned host buffers (pipelines/base_pipeline.py:114-120 pattern)"},
            "gpu_launches": launches_per_step * K, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "train_step": train,
            "wall_s_timed_region": wall,
        }
        _emit(line)
    if dist is not None:
        dist.destroy_process_group()


def _ncu_traffic(engine):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture
    (profiles/r1x_traffic.json, written from profiles/r1x_tc_fine_kernel_ncu.txt); None if there is no capture for this engine."""
    p = ROOT / "profiles" / "r1x_traffic.json"
    if not p.exists():
        return None
    d = json.loads(p.read_text()).get(engine)
    return d


def _engine_is_tc(lib, ccfg):
    import ctypes as C
    from nrhints_b200 import _lib
    c2 = _lib.NrhConfig.from_buffer_copy(ccfg)
    c2.mlp_impl = _lib.NRH_MLP_TCGEN05
    return lib.nrh_check_config(C.byref(c2)) == 0


if __name__ == "__main__":
    main()
