#!/usr/bin/env python
"""Headline benchmark of the NRHints ray-march hot path (BASELINE.json config #2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rays R] [--mlp auto|fp32|tcgen05]

A step = one NeuSHintRenderer.forward over 4096 rays x 128 samples (64 coarse + 64 importance, one shadow
ray of 64 + 64 samples per primary ray, both hints, white background, inference mode) of the synthetic
800x800 workload (nrhints_b200/workload.py), random-init (geometric) weights at seed 3407.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.

--impl reference times the UNMODIFIED reference (iamNCJ/NRHints, installed byte-for-byte into baseline/_ref/ by
baseline/install_ref.py; it travels to the GPU box with the snapshot) on the box's host cores: its own
NeuSHintRenderer.forward on the same weights and rays, one 512-ray chunk (the reference's inference_chunk_size) of the
4096-ray workload per step, all host threads.  Only if baseline/_ref is absent does it fall back to the oracle port.
The main arm also reports `gpu_baseline` (the same reference module on this B200 through PyTorch CUDA, TF32 off) and
`parity_vs_reference` (our output against the reference's on the very rays that were timed).
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
# backward of a training step per fine point: SDF phase A (918 016) + phase B incl. the feature head (1 049 088)
# + 17 weight-gradient products (2 x 918 016 + 131 072) + reflectance adjoint chain and weight gradients (2 x 579 584)
FLOP_PER_FINE_POINT_BWD = 918016.0 + 1049088.0 + (2 * 918016.0 + 131072.0) + 2 * 579584.0
FLOP_PER_RAY_TRAIN = FLOP_PER_RAY + 128 * FLOP_PER_FINE_POINT_BWD


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
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        sm, pw, mx, reasons = [], [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort(); pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "sm_mhz_min": sm[0] if sm else None, "power_w_median": pw[len(pw) // 2] if pw else None, "power_w_max": pw[-1] if pw else None}


def workload_config(R: int, world: int, scaling: str = "weak") -> dict:
    """The `config` object of the JSON line -- identical for the main arm and the reference arm."""
    return {"workload": "NRHints NeuSHintRenderer.forward, BASELINE config #2: 4096 rays x (64+64) samples, shadow ray 64+64, "
                        "shadow+specular hints, white bg, 800x800 synthetic scene, geometric-init weights seed 3407",
            "rays_per_gpu": R, "samples_per_ray": 128, "parallelism": f"rays sharded x{world}, no data-path collective",
            "l2": "256 MiB buffer rewritten between timed steps (L2 flush); per-step working set ~0.9 GB > 126 MB L2"}


METRIC = "rays/sec (4096 rays x 128 samples forward render)"
REF_CHUNK = 512                      # the reference's inference_chunk_size (models/neus_hint_model.py:212)


def _reference_module(device="cpu", state=None):
    """The unmodified reference NeuSHintRenderer (baseline/_ref) with the benchmark's weights: torch.manual_seed(3407) + geometric
    init gives a state_dict bit-identical to ours (checked when the fixtures are generated), loaded explicitly anyway."""
    sys.path.insert(0, str(ROOT / "baseline"))
    import ref_loader
    if not ref_loader.available():
        return None, None
    ns = ref_loader.load()
    import nrhints_b200 as nb
    if state is None:
        torch.manual_seed(3407)
        state = nb.NeuSHintRenderer(nb.NeuSModelConfig()).state_dict()
    torch.manual_seed(3407)
    ref = ns.NeuSHintRenderer(ns.NeuSModelConfig())
    ref.load_state_dict({k: v.detach().clone().cpu() for k, v in state.items()}, strict=True)
    return ref.to(device), ns


def _ref_forward_chunks(ref, ns, rays: dict, device, chunk=REF_CHUNK):
    """reference forward over `rays` in chunks, as its own evaluation loop does (pipelines/base_pipeline.py:112-120)."""
    outs = []
    bg = torch.ones(1, 3, device=device)
    n = rays["origins"].shape[0]
    for i0 in range(0, n, chunk):
        b = ns.RayBundle(**{k: v[i0:i0 + chunk].to(device) for k, v in rays.items()})
        with torch.no_grad():
            o = ref.forward(b, is_training=False, background_rgb=bg)
        outs.append({"rgb": o.rgb.detach(), "depth": o.depth.detach(), "visibilities": o.visibilities.detach(),
                     "weights": o.weights.detach(), "normals": o.normalized_analytic_normals.detach()})
    return {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}


def cpu_reference_leg(n_chunks: int, warm: bool = True):
    """`n_chunks` 512-ray chunks of the 4096-ray workload through the reference on the host cores -> (cpu_baseline dict, seconds per
    chunk list).  Falls back to the oracle port when baseline/_ref is absent."""
    from nrhints_b200.workload import synthetic_rays
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rays = synthetic_rays(4096, seed=3407)
    ref, ns = _reference_module("cpu")
    kind = "reference"
    if ref is None:
        kind = "port"
        from oracle import nrh_oracle as orc
        import nrhints_b200 as nb
        cfg = nb.NeuSModelConfig()
        torch.manual_seed(3407)
        state = {k: v.detach().clone() for k, v in nb.NeuSHintRenderer(cfg).state_dict().items()}
        ocfg = orc.OracleConfig.from_model_config(cfg)

    def one(ci):
        sl = {k: v[ci * REF_CHUNK:(ci + 1) * REF_CHUNK] for k, v in rays.items()}
        t0 = time.perf_counter()
        if kind == "reference":
            _ref_forward_chunks(ref, ns, sl, "cpu")
        else:
            with torch.no_grad():
                orc.render_forward(state, ocfg, sl["origins"], sl["directions"], sl["pl_positions"], sl["nears"], sl["fars"],
                                   background_rgb=torch.ones(1, 3))
        return time.perf_counter() - t0
    if warm:
        small = {k: v[:64] for k, v in rays.items()}
        if kind == "reference":
            _ref_forward_chunks(ref, ns, small, "cpu")
    times = [one(ci % 8) for ci in range(n_chunks)]
    mean = sum(times) / len(times)
    base = {"value": REF_CHUNK / mean, "unit": "rays/s", "cores": cores, "kind": kind, "host_cores": os.cpu_count(),
            "sample": f"{n_chunks} chunk(s) of {REF_CHUNK} rays x 128 samples of the same 4096-ray workload (the reference's "
                      f"inference_chunk_size; a 4096-ray batch needs ~58 GB), torch.set_num_threads({cores}); "
                      + ("UNMODIFIED reference NeuSHintRenderer.forward from baseline/_ref (manifest-verified)" if kind == "reference"
                         else "oracle port (baseline/_ref absent)")}
    return base, times


def gpu_reference_leg(model, dev, rays: dict, flush):
    """the unmodified reference module on THIS B200 through PyTorch CUDA (TF32 off, the torch default),
    same weights and rays -- as 8 chunks of 512 (its own evaluation loop) and un-chunked -- and our output against its output."""
    import nrhints_b200 as nb
    ref, ns = _reference_module(dev, state=model.state_dict())          # the weights `model` holds NOW (the training-step
    if ref is None:                                                     # measurement above has moved them a few Adam steps)
        return None, None
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    drays = {k: v.to(dev) for k, v in rays.items()}
    n = drays["origins"].shape[0]

    def timed(chunk, reps=3):
        with torch.device(dev):                     # the reference has device-less tensor constructors (models/neus_hint_model.py:327);
            out = _ref_forward_chunks(ref, ns, drays, dev, chunk)      # its trainer sets the default tensor type to CUDA (trainer/trainer.py:50)
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                out = _ref_forward_chunks(ref, ns, drays, dev, chunk)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2], out
    ms_chunked, out = timed(REF_CHUNK)
    ms_full = None
    try:
        ms_full, _ = timed(n)
    except torch.OutOfMemoryError:
        pass
    torch.cuda.empty_cache()
    with torch.no_grad():
        mine = model(nb.RayBundle(**drays), is_training=False, background_rgb=torch.ones(1, 3, device=dev))
    mse = float(((mine.rgb - out["rgb"]).double() ** 2).mean())
    d_rgb_ray = (mine.rgb - out["rgb"]).abs().amax(dim=-1)
    d_vis_ray = (mine.visibilities - out["visibilities"]).abs().reshape(-1)
    parity = {"rays": n, "max_abs_d_rgb": float((mine.rgb - out["rgb"]).abs().max()),
              # the maxima below are set by single rays whose far-end importance sample lands in another bin under fp32 noise (the
              # reference's own CPU and GPU runs differ on those too, tests/nrh_testlib.compare_outputs): how many rays that is
              "rays_with_d_rgb_over_1e-3": int((d_rgb_ray > 1e-3).sum()), "rays_with_d_visibility_over_1e-3": int((d_vis_ray > 1e-3).sum()),
              "median_abs_d_rgb": float(d_rgb_ray.median()), "median_abs_d_visibility": float(d_vis_ray.median()),
              "psnr_between_db": float(10 * math.log10(1.0 / max(mse, 1e-30))),
              "max_abs_d_depth": float((mine.depth - out["depth"]).abs().max()),
              "max_abs_d_visibility": float((mine.visibilities - out["visibilities"]).abs().max()),
              "gate": "BASELINE.json: per-pixel max |d rgb| < 1e-3, PSNR delta < 0.01 dB (<=> the two images are > 60 dB apart)",
              "against": "unmodified reference (baseline/_ref) on the same B200, PyTorch CUDA fp32, identical weights and rays"}
    gpu = {"value": n / (ms_chunked * 1e-3), "unit": "rays/s", "ms": ms_chunked, "kind": "reference",
           "what": f"unmodified reference NeuSHintRenderer.forward on this GPU through PyTorch CUDA (allow_tf32=False), {n} rays as "
                   f"{n // REF_CHUNK} chunks of {REF_CHUNK}, 1 warm-up + median of 3",
           "unchunked": ({"value": n / (ms_full * 1e-3), "ms": ms_full} if ms_full else None)}
    return gpu, parity


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu_reference_leg(1, warm=True)                      # warm-up: one small + one full chunk (threads, allocator), whatever --warmup says
    base, times = cpu_reference_leg(args.steps, warm=False)
    ms = 1e3 * sum(times) / len(times)
    value = REF_CHUNK / (sum(times) / len(times))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.rays, args.gpus), "rays_per_step": REF_CHUNK,
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

    fine_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for e in fine_ev:
        e.record()                                # materialise the CUDA events; the library re-records them around the fine pass
    fine_ms = []

    def step(timed: bool = False):
        return model(dev_rays, is_training=False, background_rgb=bg, _fine_events=fine_ev if timed else None)

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
    for _ in range(3):                            # the dominant kernel timed INSIDE a step (events recorded by the library)
        flush.zero_()
        step(timed=True)
        torch.cuda.synchronize()
        fine_ms.append(fine_ev[0].elapsed_time(fine_ev[1]))
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
                "kernel_ms": k_ms, "kernel_ms_in_step": sum(fine_ms) / len(fine_ms),
                "achieved_in_step": FLOP_PER_FINE_POINT * R * 128 / (sum(fine_ms) / len(fine_ms) * 1e-3) / 1e12,
                "flop_per_launch": FLOP_PER_FINE_POINT * R * 128, "traffic": (_ncu_traffic(engine) or {}).get("bytes_per_launch"), "traffic_unit": "B per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                "traffic_detail": _ncu_traffic(engine),
                "whole_step": {"achieved_tflops": FLOP_PER_RAY * R / (ms_per_step * 1e-3) / 1e12,
                               "frac_of_tensor_peak": FLOP_PER_RAY * R / (ms_per_step * 1e-3) / 1e12 / peak_sus,
                               "hbm_algorithmic_gbs": BYTES_PER_RAY * R / (ms_per_step * 1e-3) / 1e9,
                               "frac_of_hbm_peak": BYTES_PER_RAY * R / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"]}}

    # ---- sustained clocks / board power under this workload: the forward looped for ~2.5 s while nvidia-smi samples every 20 ms.
    # The tcgen05 kernels run into the board's power cap (sw_power_cap): the SM clock settles well below its maximum, so the step
    # time is set by energy per point, not by issue slots.  Reported, never used for `value`.
    sustained = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sp = ClockSampler(local_rank)
        sp.start()
        t_end, n_it = time.perf_counter() + 2.5, 0
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        while time.perf_counter() < t_end:
            for _ in range(10):
                step()
            n_it += 10
            torch.cuda.synchronize()
        b.record(); torch.cuda.synchronize()
        sustained = sp.stop()
        sustained["ms_per_step_back_to_back"] = a.elapsed_time(b) / n_it
        sustained["what"] = "forward steps back to back for 2.5 s (no L2 flush in between), nvidia-smi every 20 ms"

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
        opt = pipe.make_optimizer(capturable=True)               # parameters / gradients / moments re-homed into flat buffers; step count
                                                                 # and learning rate on the device so that the step can live in a CUDA graph

        sync = (lambda: allreduce_flat(opt.flat_grads())) if dist is not None else None

        def train_step():
            # NRHintPipeline.train_step: ray generation -> nrh_render_train_forward -> nrh_train_loss -> nrh_render_backward (gradients
            # written straight into FlatAdam's flat buffer) -> the one collective of the loop on that buffer (its mean folded into the
            # Adam launch) -> nrh_adam_step; no autograd graph, no library GEMM
            pipe.train_step(pixels, global_step=60000, optimizer=opt, grad_sync=sync)

        def train_step_autograd():
            # the same step the way an unmodified trainer drives it (trainer/trainer.py:269-283): loss.backward() lands in ONE autograd node
            opt.zero_grad()
            res = pipe(pixels, global_step=60000)
            loss = pipe.get_train_loss_dict(res, pixels)["loss"]
            loss.backward()
            scale = allreduce_flat(opt.flat_grads()) if dist is not None else 1.0
            opt.step(grad_scale=scale)

        def time_train(fn, n):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            tev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            if dist is not None:
                dist.barrier()
            for a, b in tev:
                flush.zero_()
                a.record(); fn(); b.record()
            torch.cuda.synchronize()
            tt = torch.tensor([sum(a.elapsed_time(b) for a, b in tev) / n], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        tsteps = max(3, min(K, 5))
        t_ms = time_train(train_step, tsteps)
        train_launches = int(model.last_launch_count + getattr(model, "last_backward_launch_count", 0))
        t_auto_ms = time_train(train_step_autograd, tsteps)
        t_graph_ms = None
        try:                                                   # the whole step (weight pack ... Adam, all-reduce included) as ONE CUDA graph
            graphed = pipe.capture_train_step(pixels, opt, grad_sync=sync, global_step=60000)
            t_graph_ms = time_train(lambda: graphed(pixels, 60000), tsteps)
        except Exception as e:                                 # capture is an optimisation, never a requirement
            print("CUDA graph capture of the training step failed:", repr(e), file=sys.stderr)
        ar_ms = None
        if dist is not None:                                   # the step's one collective, timed alone on the compute stream
            for _ in range(3):
                allreduce_flat(opt.flat_grads())
            torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                allreduce_flat(opt.flat_grads())
            b.record(); torch.cuda.synchronize()
            ar = torch.tensor([a.elapsed_time(b) / 10], device=dev, dtype=torch.float64)
            dist.all_reduce(ar, op=dist.ReduceOp.MAX)
            ar_ms = float(ar.item())
        tf = FLOP_PER_RAY_TRAIN * R / (t_ms * 1e-3) / 1e12
        peaks = measured_peaks()
        train = {"value": world * R / (t_ms * 1e-3), "unit": "rays/s", "ms_per_step": t_ms, "steps": tsteps,
                 "ms_per_step_through_autograd_node": t_auto_ms, "ms_per_step_cuda_graph": t_graph_ms,
                 "library_launches_forward_plus_backward": train_launches,
                 "allreduce_ms": ar_ms, "allreduce_bytes": sum(g.numel() * 4 for g in opt.flat_grads()),
                 "allreduce_share_of_step": (ar_ms / t_ms if ar_ms else None),
                 "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops_sustained"] if "bf16_tflops_sustained" in peaks else None,
                              "unit": "TFLOP/s", "frac": (tf / peaks["bf16_tflops_sustained"]) if "bf16_tflops_sustained" in peaks else None,
                              "flop_per_step": FLOP_PER_RAY_TRAIN * R,
                              "note": "logical FLOPs of forward + backward (each SDF product is issued as 3 fp16 MMAs) over the whole step, "
                                      "against the measured sustained cuBLAS bf16 peak"},
                 "what": "BASELINE config #3: ray generation + forward + L1/eikonal loss + backward + Adam on 4096 rays/GPU (the reference's train_iter), is_training=True "
                         "(jitter, global_step 60000) as NRHintPipeline.train_step: nrh_render_train_forward + nrh_train_loss + nrh_render_backward + nrh_adam_step -- every "
                         "GEMM (SDF forward-with-tape, second-order backward, reflectance forward / backward, all weight-gradient reductions) on hand-written tcgen05 "
                         "kernels, gradients written straight into the flat all-reduce buffer, no autograd graph, no library GEMM"
                         + ("; flat-buffer gradient all-reduce" if dist is not None else "")}
        torch.set_grad_enabled(False)
        del opt

    # ---- strong scaling (the reference's semantics: a FIXED global batch split over the ranks, trainer/trainer.py:118) ----------
    # R rays in total => R / world per GPU; the forward of a rank is captured in a CUDA graph (at 512 rays per GPU the ~35 launches
    # of a step are launch-latency bound otherwise).  Same barrier + device-event + max-over-ranks timing as the headline.
    strong = None
    if world > 1 and R % world == 0:
        Rs = R // world
        full = synthetic_rays(R, seed=3407)
        mine_rays = nb.RayBundle(**{k: v[rank * Rs:(rank + 1) * Rs] for k, v in full.items()}).to(dev)

        def time_steps(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            for a, b in ev:
                flush.zero_()
                a.record(); fn(); b.record()
            torch.cuda.synchronize()
            tt = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / K], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        eager_ms = time_steps(lambda: model(mine_rays, is_training=False, background_rgb=bg))
        graph_ms = None
        try:
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                model(mine_rays, is_training=False, background_rgb=bg)
            torch.cuda.current_stream(dev).wait_stream(side)
            with torch.cuda.graph(g):
                model(mine_rays, is_training=False, background_rgb=bg)
            graph_ms = time_steps(g.replay)
        except Exception as e:                                  # capture is an optimisation, never a requirement
            print("CUDA graph capture of the forward failed:", repr(e), file=sys.stderr)
        best = min(x for x in (eager_ms, graph_ms) if x is not None)
        strong = {"scaling": "strong", "global_rays": R, "rays_per_gpu": Rs, "value": R / (best * 1e-3), "unit": "rays/s",
                  "ms_per_step": best, "ms_per_step_eager": eager_ms, "ms_per_step_cuda_graph": graph_ms,
                  "what": "the same 4096-ray batch split over the ranks (no collective in the forward)"}
        if not args.no_train:
            # the training step the way the reference shards it (trainer/trainer.py:118: batch_size // world_size rays per rank), gradient
            # all-reduce on the flat buffer included, eager and as one CUDA graph per rank
            try:
                torch.set_grad_enabled(True)
                from types import SimpleNamespace
                from nrhints_b200.grad_sync import allreduce_flat
                from nrhints_b200.workload import synthetic_pixel_bundle
                pb_s, cam_s = synthetic_pixel_bundle(R, seed=3407)
                px_s = SimpleNamespace(**{k: (v[rank * Rs:(rank + 1) * Rs].to(dev) if isinstance(v, torch.Tensor) and v.shape[:1] == (R,) else v)
                                          for k, v in vars(pb_s).items()})
                pipe_s = nb.NRHintPipeline(cfg, nb.RayGeneratorConfig(), nb.CameraModel(**cam_s), 64, mlp_impl=args.mlp)
                pipe_s.renderer = model
                pipe_s = pipe_s.to(dev)
                opt_s = pipe_s.make_optimizer(capturable=True)
                sync_s = lambda: allreduce_flat(opt_s.flat_grads())      # noqa: E731
                t_eager = time_steps(lambda: pipe_s.train_step(px_s, global_step=60000, optimizer=opt_s, grad_sync=sync_s))
                t_graph = None
                try:
                    gs = pipe_s.capture_train_step(px_s, opt_s, grad_sync=sync_s, global_step=60000)
                    t_graph = time_steps(lambda: gs(px_s, 60000))
                except Exception as e:
                    print("CUDA graph capture of the sharded training step failed:", repr(e), file=sys.stderr)
                tb = min(x for x in (t_eager, t_graph) if x is not None)
                strong["train_step"] = {"value": R / (tb * 1e-3), "unit": "rays/s", "ms_per_step": tb, "ms_per_step_eager": t_eager,
                                        "ms_per_step_cuda_graph": t_graph, "rays_per_gpu": Rs,
                                        "what": "config #3 with the reference's batch split: 4096 rays per step in total, gradient all-reduce on the flat buffer inside the step"}
                torch.set_grad_enabled(False)
            except Exception as e:
                print("strong-scaling training step failed:", repr(e), file=sys.stderr)

    cpu = gpu_base = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference_leg(2, warm=True)           # 2 x 512 rays through the unmodified reference on the host cores
        gpu_base, parity = gpu_reference_leg(model, dev, synthetic_rays(R, seed=3407 + rank), flush)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if engine == "fp32-simt" else "f32 results; SDF network: fp16 hi/lo split operands x3 MMAs, fp32 accumulate; "
                                                         "reflectance network: SINGLE-PASS fp16 operands, fp32 accumulate (narrower than the reference's fp32)",
            "data": "synthetic", "config": workload_config(R, world), "engine": engine,
            "e2e": {"value": world * R / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "pinned host RayBundle -> device, forward, full RenderOutput -> pinned host buffers (pipelines/base_pipeline.py:114-120 pattern)"},
            "gpu_launches": launches_per_step * K, "clocks": clocks, "clocks_sustained": sustained, "roofline": roof, "cpu_baseline": cpu, "gpu_baseline": gpu_base, "parity_vs_reference": parity,
            "train_step": train, "strong_scaling": strong,
            "wall_s_timed_region": wall,
        }
        _emit(line)
    if dist is not None:
        # leave without tearing NCCL down object by object: every rank has printed / finished, a last barrier lines them up, and the
        # process exits at once (communicator destructors racing CUDA-graph and allocator teardown can stall for minutes)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def _ncu_traffic(engine):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture
    (profiles/r1x_traffic.json, written from profiles/r1x_tc_fine_kernel_ncu.txt); None if there is no capture for this engine."""
    p = ROOT / "profiles" / "r2d_traffic.json"
    if not p.exists():
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
