"""Secondary measurements (not the headline bench.py line): hash encoding, outside-NeRF render, training step of the
interim autograd backend, grid SDF query.  One JSON line per measurement; CUDA events, L2 flushed between iterations.

    python tools_bench_extra.py [--what hash,outside,train,grid,torchgpu] [--iters 5]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import nrhints_b200 as nb                                   # noqa: E402
from nrhints_b200.workload import synthetic_rays            # noqa: E402


def timed(fn, iters, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="hash,outside,train,grid,torchgpu,image")
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    what = set(args.what.split(","))
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}

    if "hash" in what:
        from nrhints_b200.encodings import HashEncoding
        torch.manual_seed(0)
        enc = HashEncoding().to(dev)
        N = 4096 * 128
        pts = torch.rand(N, 3, device=dev)
        with torch.no_grad():
            ms = timed(lambda: enc(pts), args.iters, flush)
        alg = N * (12 + 128)                                   # 12 B in + 32 fp32 out per point
        gather = N * 16 * 8 * 8                                # table bytes touched (L2-resident 64 MiB table)
        print(json.dumps({"what": "nrh_hash_encode forward", "points": N, "ms": ms, "Mpoints_per_s": N / ms / 1e3,
                          "algorithmic_hbm_gbs": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / peaks["hbm_gbs"],
                          "table_gather_gbs": gather / ms / 1e6, "bound": "L2 gathers (64 MiB table resident in the 126 MB L2)"}))
        out = enc(pts)
        g = torch.randn_like(out)
        def bwd():
            enc.hash_table.grad = None
            enc(pts).backward(g)
        ms = timed(bwd, args.iters, flush)
        print(json.dumps({"what": "nrh_hash_encode forward+backward (table gradient, fp32 atomics)", "points": N, "ms": ms}))

    if "outside" in what:
        torch.manual_seed(3407)
        cfg = nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(use_outside_nerf=True))
        m = nb.NeuSHintRenderer(cfg).to(dev)
        R = 4096
        rays = nb.RayBundle(**synthetic_rays(R, seed=3407)).to(dev)
        bg = torch.ones(1, 3, device=dev)
        with torch.no_grad():
            ms = timed(lambda: m(rays, background_rgb=bg), args.iters, flush)
        print(json.dumps({"what": "forward render with the outside NeRF (4096 rays x (128 + 32 outside) samples)", "ms": ms,
                          "rays_per_s": R / ms * 1e3, "launches": m.last_launch_count}))

    if "train" in what:
        torch.manual_seed(3407)
        cfg = nb.NeuSModelConfig()
        m = nb.NeuSHintRenderer(cfg).to(dev)
        opt = torch.optim.Adam(m.parameters(), lr=5e-4)
        for R in (512, 4096):
            rays = nb.RayBundle(**synthetic_rays(R, seed=3407)).to(dev)
            bg = torch.ones(1, 3, device=dev)
            gt = torch.rand(R, 3, device=dev)
            def step():
                opt.zero_grad(set_to_none=True)
                out = m(rays, is_training=True, background_rgb=bg, global_step=60000)
                rgb_loss = torch.nn.functional.l1_loss(out.rgb, gt, reduction="sum") / (R + 1e-5)
                gerr = (torch.linalg.norm(out.analytic_normals, ord=2, dim=-1) - 1.0) ** 2
                eik = (out.relax_inside_sphere * gerr).sum() / (out.relax_inside_sphere.sum() + 1e-5)
                (rgb_loss + 0.1 * eik).backward()
                opt.step()
            torch.cuda.reset_peak_memory_stats()
            ms = timed(step, max(3, args.iters // 2), flush)
            print(json.dumps({"what": "training step (BASELINE config #3: fwd + bwd + Adam), interim autograd backend", "rays": R,
                              "ms": ms, "rays_per_s": R / ms * 1e3, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))

    if "torchgpu" in what:
        # the reference ALGORITHM as unfused PyTorch ops on the same B200 (oracle port, fp32,
        # allow_tf32 = False as in the reference), chunked at the reference's inference_chunk_size = 512 and at 4096
        from oracle import nrh_oracle as orc
        torch.backends.cuda.matmul.allow_tf32 = False
        cfg = nb.NeuSModelConfig()
        torch.manual_seed(3407)
        state = {k: v.detach().to(dev) for k, v in nb.NeuSHintRenderer(cfg).state_dict().items()}
        ocfg = orc.OracleConfig.from_model_config(cfg)
        W = orc.effective_weights(state)
        R = 4096
        rays = {k: v.to(dev) for k, v in synthetic_rays(R, seed=3407).items()}
        bg = torch.ones(1, 3, device=dev)
        for chunk in (512, 4096):
            def run():
                for i in range(0, R, chunk):
                    with torch.no_grad():
                        orc.render_forward(W, ocfg, rays["origins"][i:i + chunk], rays["directions"][i:i + chunk],
                                           rays["pl_positions"][i:i + chunk], rays["nears"][i:i + chunk], rays["fars"][i:i + chunk],
                                           background_rgb=bg, effective=True)
            torch.cuda.reset_peak_memory_stats()
            ms = timed(run, 3, flush)
            print(json.dumps({"what": "reference algorithm as unfused PyTorch CUDA ops on the same B200 (oracle port, fp32, TF32 off)",
                              "rays": R, "chunk": chunk, "ms": ms, "rays_per_s": R / ms * 1e3,
                              "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))

    if "image" in what:
        # one 800x800 view (640 000 rays) from pinned host rays to pinned host maps
        import time
        torch.manual_seed(3407)
        m = nb.NeuSHintRenderer(nb.NeuSModelConfig()).to(dev)
        N = 800 * 800
        rays = nb.RayBundle(**synthetic_rays(N, seed=1)).pin_memory()
        bg = torch.ones(1, 3)
        for chunk in (4096, 16384, 65536):
            m.render_image(rays, background_rgb=bg, chunk_rays=chunk)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = m.render_image(rays, background_rgb=bg, chunk_rays=chunk)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(json.dumps({"what": "render_image: 800x800 view, pinned host rays -> pinned host per-ray maps", "chunk_rays": chunk,
                              "s": dt, "rays_per_s": N / dt, "d2h_bytes": sum(v.numel() * 4 for v in out.values()),
                              "workspace_gb": m._workspace.numel() / 2**30}))

    if "grid" in what:
        torch.manual_seed(3407)
        m = nb.NeuSHintRenderer(nb.NeuSModelConfig()).to(dev)
        N = 256 ** 3
        pts = (torch.rand(N, 3, device=dev) - 0.5) * 2.0
        with torch.no_grad():
            ms = timed(lambda: m.sdf_query(pts), 3, flush)
        flop = 918016.0 * N
        print(json.dumps({"what": "nrh_sdf_query (sdf only) over a 256^3 grid (extract_fields)", "points": N, "ms": ms,
                          "Mpoints_per_s": N / ms / 1e3, "logical_tflops": flop / ms / 1e9}))


if __name__ == "__main__":
    main()
