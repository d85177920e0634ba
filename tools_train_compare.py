#!/usr/bin/env python
"""BASELINE.json configs #4 / #5 at test scale: N optimisation steps of the reference's training loop (trainer/trainer.py:269-283:
next batch -> pipeline.forward -> loss dict -> zero_grad / backward / Adam step / LR scheduler) on a synthetic scene, run three ways
with identical seeds, batches and hyper-parameters on the same GPU:

  ref     the UNMODIFIED reference (baseline/_ref): its BaseNRHintPipeline, its renderer, torch.optim.Adam
  dropin  the same reference pipeline object with ONE line changed -- `pipeline.renderer = nrhints_b200.NeuSHintRenderer(cfg.model)`
           -- still torch.optim.Adam and the reference's loss / ray generator
  native  nrhints_b200.NRHintPipeline: CUDA ray generation, fused renderer, fused loss, FlatAdam (the full B200-native step)

and reports loss curves, steps/s and the PSNR of held-out views rendered at the end (same evaluation code for all arms).

The scene: the reference network with perturbed, sharpened weights ("teacher": tests/nrh_testlib.make_state('sharp')) rendered by
the reference renderer itself from `--views` cameras on the radius-4 sphere (64x64 pixels each, one point light per view at radius
4.5).  The real "Cat" capture of scripts/train_real.sh is not available offline; as in its nr-hints-cam-opt preset the camera poses
carry noise (RayGeneratorConfig.cam_position_noise_std / cam_orientation_noise_std) that the SO3xR3 refinement has to absorb.
The schedule constants are shortened in proportion to the run (warm_up_end, anneal_end, end_iter), identically for every arm.

    python tools_train_compare.py --steps 1000 --out profiles/r2_train_compare.json

Data parallel (the reference's DDP semantics, trainer/trainer.py:88-93,118: the SAME global batch split `batch // world_size` per
rank, gradients averaged by one all-reduce per step): launch the `native` arm under torchrun --

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools_train_compare.py --steps 1000 --arms native --out profiles/r2_train_compare_2gpu.json

every rank renders the ground truth (deterministic), draws the identical batch sequence and takes its contiguous slice; the step is
NRHintPipeline.train_step with grad_sync.allreduce_flat on FlatAdam's flat gradient buffers; rank 0 evaluates and reports.
"""
import argparse
import json
import math
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "baseline")); sys.path.insert(0, str(ROOT / "tests"))


def pose(theta, phi, radius=4.0):
    c = torch.tensor([radius * math.cos(phi) * math.sin(theta), radius * math.sin(phi), radius * math.cos(phi) * math.cos(theta)])
    fwd = -c / c.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 1.0, 0.0])); right = right / right.norm()
    up = torch.linalg.cross(right, fwd)
    m = torch.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, -fwd, c
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--views", type=int, default=14)
    ap.add_argument("--res", type=int, default=64)
    ap.add_argument("--arms", default="ref,dropin,native")
    ap.add_argument("--preset", default="NRHintsCamOpt", choices=["NRHints", "NRHintsCamOpt"])
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import os
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        assert args.arms == "native", "only the native arm is data parallel (the reference arms need its own DDP trainer)"
        assert args.batch % world == 0
    dev = torch.device("cuda", torch.cuda.current_device())
    import ref_loader
    import nrh_testlib as T
    import nrhints_b200 as nb
    ns = ref_loader.load_pipeline()
    M, C, RG = ns.model, ns.configs, ns.ray_generator
    from data.shm_helper import NRDataSHMInfo

    H = W = args.res
    n_views, n_train = args.views, args.views - 2
    fx = 0.5 * W / math.tan(0.5 * 0.6911)
    cam = dict(H=H, W=W, cx=W / 2.0, cy=H / 2.0, fx=fx, fy=fx, zn=2.0, zf=6.0)
    g = torch.Generator().manual_seed(2024)
    poses = torch.stack([pose(2 * math.pi * i / n_views + 0.1, 0.35 + 0.25 * math.sin(1.7 * i)) for i in range(n_views)])
    pls = 4.5 * torch.nn.functional.normalize(torch.randn(n_views, 3, generator=g) + torch.tensor([0.0, 0.8, 0.0]), dim=-1)
    S = args.steps
    model_kw = dict(batch_size=args.batch, warm_up_end=max(S // 10, 1), anneal_end=max(S // 2, 1), end_iter=S)
    rg_kw = dict(cam_opt_mode="SO3xR3", cam_position_noise_std=0.01, cam_orientation_noise_std=0.005) if args.preset == "NRHintsCamOpt" else {}

    # ---- ground truth: the teacher network rendered by the reference renderer (noise-free poses) -------------------------------
    torch.manual_seed(3407)
    teacher = M.NeuSHintRenderer(M.NeuSModelConfig())
    teacher.load_state_dict(T.make_state("sharp", nb.NeuSModelConfig()), strict=True)
    teacher = teacher.to(dev)
    clean_rg = RG.RayGenerator(ns.camera_model.CameraModel(**cam), n_views, RG.RayGeneratorConfig()).to(dev)
    ww, hh = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="xy")
    imgs = torch.empty(n_views, H, W, 3)
    t0 = time.perf_counter()
    with torch.no_grad(), torch.device(dev):
        for v in range(n_views):
            pb = ns.data_loader.RawPixelBundle(img_indices=None, h_indices=hh.reshape(-1, 1).to(dev), w_indices=ww.reshape(-1, 1).to(dev),
                                               poses=poses[v].to(dev)[None].repeat(H * W, 1, 1), pls=pls[v].to(dev)[None].repeat(H * W, 1), rgb_gt=None)
            rays = clean_rg(pb)
            out = []
            for i0 in range(0, H * W, 512):
                out.append(teacher.forward(rays[i0:i0 + 512], is_training=False, background_rgb=torch.ones(1, 3)).rgb.detach())
            imgs[v] = torch.cat(out).reshape(H, W, 3).cpu()
    print(f"ground truth: {n_views} views of {H}x{W} in {time.perf_counter() - t0:.1f} s", file=sys.stderr)
    del teacher

    def batches():
        """the reference's PixelSampler with the ALL_IMAGES strategy (data/data_loader.py:57-76; trainer/trainer.py:118-125)"""
        image_rng, pixel_rng = np.random.default_rng(3407), np.random.default_rng(3407)
        while True:
            ii = image_rng.choice(n_train, args.batch)
            hi, wi = pixel_rng.choice(H, args.batch), pixel_rng.choice(W, args.batch)
            yield dict(img_indices=torch.from_numpy(ii)[..., None], h_indices=torch.from_numpy(hi)[..., None].float(),
                       w_indices=torch.from_numpy(wi)[..., None].float(), rgb_gt=imgs[ii, hi, wi], poses=poses[ii], pls=pls[ii])

    def lr_lambda(cfg_model):
        def f(it):                                                   # trainer/trainer.py:105-111
            if it < cfg_model.warm_up_end:
                return it / cfg_model.warm_up_end
            p = (it - cfg_model.warm_up_end) / (cfg_model.end_iter - cfg_model.warm_up_end)
            return (np.cos(np.pi * p) + 1.0) * 0.5 * (1 - cfg_model.lr_alpha) + cfg_model.lr_alpha
        return f

    def build(arm):
        torch.manual_seed(3407); torch.cuda.manual_seed(3407)
        if arm in ("ref", "dropin"):
            cfg = getattr(C, args.preset)(model=M.NeuSModelConfig(**model_kw), ray_generator=RG.RayGeneratorConfig(**rg_kw))
            shm = NRDataSHMInfo(total_image_num=n_views, num_image_per_split=[n_train, 0, 2], camera=ns.camera_model.CameraModel(**cam),
                                imgs_shm_name="", poses_shm_name="", pls_shm_name="")
            pipe = ns.pipeline.BaseNRHintPipeline(cfg, shm)
            if arm == "dropin":
                sd = pipe.renderer.state_dict()
                pipe.renderer = nb.NeuSHintRenderer(cfg.model)
                pipe.renderer.load_state_dict(sd, strict=True)
            pipe = pipe.to(dev)
            opt = torch.optim.Adam(pipe.get_param_groups())
            mk = lambda b: ns.data_loader.RawPixelBundle(**b).to(dev)        # noqa: E731
            return pipe, opt, cfg.model, mk
        cfg_model = nb.NeuSModelConfig(**model_kw)
        pipe = nb.NRHintPipeline(cfg_model, nb.RayGeneratorConfig(**rg_kw), nb.CameraModel(**cam), n_views).to(dev)
        opt = pipe.make_optimizer()
        from types import SimpleNamespace
        mk = lambda b: SimpleNamespace(**{k: v.to(dev) for k, v in b.items()})      # noqa: E731
        return pipe, opt, cfg_model, mk

    def evaluate(pipe):
        """held-out views (image indices n_train, n_train+1): same code for every arm -- the pipeline's ray generator (noisy pose
        of that image + its learned correction, as in get_eval_dicts without the 500 registration steps) + renderer, 512-ray chunks"""
        ps = []
        with torch.no_grad(), torch.device(dev):
            for v in (n_train, n_train + 1):
                fields = dict(img_indices=torch.full((H * W, 1), v, device=dev), h_indices=hh.reshape(-1, 1).to(dev),
                              w_indices=ww.reshape(-1, 1).to(dev), poses=poses[v].to(dev)[None].repeat(H * W, 1, 1),
                              pls=pls[v].to(dev)[None].repeat(H * W, 1), rgb_gt=imgs[v].reshape(-1, 3).to(dev))
                rgb = []
                for i0 in range(0, H * W, 512):
                    sl = {k: t[i0:i0 + 512] for k, t in fields.items()}
                    pb = ns.data_loader.RawPixelBundle(**sl) if not isinstance(pipe, nb.NRHintPipeline) else __import__("types").SimpleNamespace(**sl)
                    rays = pipe.ray_generator(pb)
                    rgb.append(pipe.renderer(rays, is_training=False, background_rgb=torch.ones(1, 3, device=dev)).rgb.detach())
                mse = float(((torch.cat(rgb) - fields["rgb_gt"]) ** 2).mean())
                ps.append(10 * math.log10(1.0 / mse))
        return ps

    results = {}
    for arm in args.arms.split(","):
        pipe, opt, cfg_model, mk = build(arm)
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda(cfg_model))
        gen = batches()
        losses, psnrs = [], []
        psnr0 = evaluate(pipe)
        torch.manual_seed(17 + rank); torch.cuda.manual_seed(17 + rank)
        sync = None
        if world > 1:
            from nrhints_b200.grad_sync import allreduce_flat
            sync = lambda: allreduce_flat(opt.flat_grads())      # noqa: E731
            per = args.batch // world
        t_start = None
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for step in range(S):
            if step == S // 10:
                torch.cuda.synchronize(); ev_a.record()
            if world > 1:                                   # this rank's contiguous slice of the global batch; fused step + one all-reduce
                pb = mk({k: v[rank * per:(rank + 1) * per] for k, v in next(gen).items()})
                with torch.device(dev):
                    ld = pipe.train_step(pb, global_step=step, optimizer=opt, grad_sync=sync)
                sched.step()
                losses.append(ld["loss"].detach())
                psnrs.append(ld["psnr"].detach())
                continue
            pb = mk(next(gen))
            with torch.device(dev):
                res = pipe(pb, global_step=step)
                ld = pipe.get_train_loss_dict(res, pb)
            opt.zero_grad()
            ld["loss"].backward()
            opt.step()
            sched.step()
            losses.append(ld["loss"].detach())
            psnrs.append(ld["psnr"] if torch.is_tensor(ld["psnr"]) else torch.tensor(ld["psnr"], device=dev))
        ev_b.record(); torch.cuda.synchronize()
        ms = ev_a.elapsed_time(ev_b) / (S - S // 10)
        Lt, Pt = torch.stack(losses).float(), torch.stack([p.float().reshape(()) for p in psnrs])
        if world > 1:                                       # report the mean over ranks (= the loss of the global batch) and the slowest rank
            dist.all_reduce(Lt); dist.all_reduce(Pt); Lt /= world; Pt /= world
            mt = torch.tensor([ms], device=dev); dist.all_reduce(mt, op=dist.ReduceOp.MAX); ms = float(mt)
        L, P = Lt.cpu().numpy(), Pt.cpu().numpy()
        k = max(S // 100, 1)
        results[arm] = {"ms_per_step": ms, "steps_per_s": 1e3 / ms, "rays_per_s": args.batch * 1e3 / ms,
                        "loss_curve": [float(L[i:i + k].mean()) for i in range(0, S, k)],
                        "train_psnr_curve": [float(P[i:i + k].mean()) for i in range(0, S, k)],
                        "final_loss_mean_last_5pct": float(L[-max(S // 20, 1):].mean()),
                        "test_psnr_before": psnr0, "test_psnr_after": evaluate(pipe)}
        extra = {n: float(p.detach().abs().max()) for n, p in pipe.ray_generator.named_parameters()}
        results[arm]["ray_generator_param_absmax"] = extra
        print(arm, {kk: (vv if not isinstance(vv, list) or len(vv) < 4 else f"[{vv[0]:.4f} .. {vv[-1]:.4f}]") for kk, vv in results[arm].items()},
              file=sys.stderr)
        del pipe, opt
        torch.cuda.empty_cache()
    if rank != 0:
        dist.barrier(); dist.destroy_process_group()
        return
    summary = {"config": {"preset": args.preset, "steps": S, "batch": args.batch, "world_size": world, "views": n_views, "res": args.res, **model_kw, **rg_kw,
                          "curve_bin_steps": max(S // 100, 1)}, "arms": results}
    if "ref" in results:
        for arm in results:
            if arm == "ref":
                continue
            a, b = np.array(results[arm]["loss_curve"]), np.array(results["ref"]["loss_curve"])
            summary.setdefault("vs_ref", {})[arm] = {
                "max_rel_loss_curve_gap": float(np.max(np.abs(a - b) / b)), "mean_rel_loss_curve_gap": float(np.mean(np.abs(a - b) / b)),
                "final_loss_ratio": results[arm]["final_loss_mean_last_5pct"] / results["ref"]["final_loss_mean_last_5pct"],
                "test_psnr_delta_db": [x - y for x, y in zip(results[arm]["test_psnr_after"], results["ref"]["test_psnr_after"])],
                "speedup_steps_per_s": results[arm]["steps_per_s"] / results["ref"]["steps_per_s"]}
    txt = json.dumps(summary)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(txt + "\n")
    print(txt)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
