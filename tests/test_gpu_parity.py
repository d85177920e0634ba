"""Parity tests proper: the CUDA path (through the C ABI, via the host mirror of NeuSHintRenderer)
against the oracle on the same seeded inputs, and against the committed reference fixtures."""
import numpy as np
import pytest
import torch

import nrh_testlib as T
import nrhints_b200 as nb
from oracle import nrh_oracle as orc

pytestmark = pytest.mark.gpu
IMPLS = ["fp32", "auto"]


def build_module(case, impl):
    """impl "auto-composed": the tcgen05 engine with gradients through the composed autograd nodes (autograd_fine.py /
    sdf_autograd.py) instead of the single fused node (fused_step.py, the default): two implementations of the same step."""
    cfg = T.make_config(case)
    sd = T.make_state(case["weights"], cfg)
    m = nb.NeuSHintRenderer(cfg, mlp_impl="auto" if impl == "auto-composed" else impl)
    m.fused_training = impl != "auto-composed"
    m.load_state_dict(sd)
    return m.cuda(), cfg, sd


def run_cuda(case, impl):
    m, cfg, sd = build_module(case, impl)
    rays, bg = T.case_inputs(case)
    bundle = nb.RayBundle(**rays).to("cuda")
    training = bool(case.get("training"))
    if training:
        torch.manual_seed(case["rng_seed"])
        torch.cuda.manual_seed(case["rng_seed"])
    with torch.no_grad():                 # pure CUDA path (with grad enabled the autograd backend would take over the fine pass)
        out = m(bundle, is_training=training, background_rgb=bg.cuda(), global_step=case.get("global_step", 0),
                return_extras=True)
    torch.cuda.synchronize()
    assert m.last_launch_count > 10
    return out, m


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", [n for n in T.CASES if not T.CASES[n].get("training")])
def test_forward_matches_oracle_and_fixture(name, impl):
    case = T.CASES[name]
    out, _ = run_cuda(case, impl)
    got = T.to_np(out)
    for k, v in got.items():
        assert np.isfinite(v).all(), k
    tol = (T.TOL if impl == "fp32" else T.TOL_TC)[case["weights"]]
    want = T.to_np(T.run_oracle(case))
    fx = np.load(T.GOLDEN_DIR / f"{name}.npz")
    ref = {k[4:]: fx[k] for k in fx.files if k.startswith("out_")}
    if name.startswith("sphere"):
        # sphere tracing is chaotic on grazing rays (the march creeps along the silhouette until the 2000-iteration cap,
        # so float noise decides the final depth -- in the reference too).  Compare the rays whose traced depth is robust,
        # judged by the oracle alone: fp32 and fp64 evaluations agree.
        w64 = T.to_np(T.run_oracle(case, dtype=torch.float64))
        robust = np.abs(want["depth"] - w64["depth"])[:, 0] < 5e-5
        assert robust.mean() >= 0.7
        got, want, ref = ({k: v[robust] for k, v in d.items()} for d in (got, want, ref))
    s1 = T.compare_outputs(got, want, label=f"cuda[{impl}]-vs-oracle[{name}]", **tol)
    s2 = T.compare_outputs(got, ref, label=f"cuda[{impl}]-vs-reference-fixture[{name}]", **tol)
    # BASELINE.json gate: PSNR delta < 0.01 dB against any ground truth <=> the two images are > 60 dB apart
    assert s1["psnr_between"] > 60 and s2["psnr_between"] > 60
    print(name, impl, {k: f"{v:.2e}" for k, v in s2.items()})


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", ["train_16x128", "train_sharp_24x128", "outside_train_8x128"])
def test_training_mode_forward(name, impl):
    """jitter on both marches (+ the outside samples with the outside NeRF) + cos annealing; the jitters are drawn by
    torch.rand on the device in the reference's order, so the oracle is fed the very same numbers."""
    case = T.CASES[name]
    m, cfg, sd = build_module(case, impl)
    rays, bg = T.case_inputs(case)
    R = case["R"]
    torch.manual_seed(7)
    jp = torch.rand([R, 1], device="cuda")
    jo = torch.rand([R, cfg.renderer.n_outside_samples], device="cuda") if cfg.renderer.use_outside_nerf else None
    js = torch.rand([R, cfg.renderer.n_shadow_samples], device="cuda")
    torch.manual_seed(7)
    with torch.no_grad():
        out = m(nb.RayBundle(**rays).to("cuda"), is_training=True, background_rgb=bg.cuda(), global_step=case["global_step"],
                return_extras=True)
    ocfg = orc.OracleConfig.from_model_config(cfg)
    with torch.no_grad():
        want = orc.render_forward(sd, ocfg, rays["origins"], rays["directions"], rays["pl_positions"], rays["nears"],
                                  rays["fars"], is_training=True, background_rgb=bg,
                                  cos_anneal=min(1.0, case["global_step"] / cfg.anneal_end),
                                  jitter_primary=jp.cpu(), jitter_shadow=js.cpu(),
                                  jitter_outside=jo.cpu() if jo is not None else None)
    got = T.to_np(out)
    assert got["weights"].shape == (R, cfg.renderer.n_samples + cfg.renderer.n_importance_samples
                                    + (cfg.renderer.n_outside_samples if cfg.renderer.use_outside_nerf else 0))
    T.compare_outputs(got, T.to_np(want), label=f"cuda[{impl}]-train[{name}]",
                      **(T.TOL if impl == "fp32" else T.TOL_TC)[case["weights"]])


@pytest.mark.parametrize("impl", IMPLS)
def test_warmup_zeroes_hints(impl):
    case = dict(T.CASES["cfg1_64x32"])
    cfg = nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(**case["renderer"]), geometry_warmup_end=100)
    m = nb.NeuSHintRenderer(cfg, mlp_impl=impl)
    m.load_state_dict(T.make_state("init", cfg)); m.cuda()
    rays, bg = T.case_inputs(case)
    with torch.no_grad():
        out = m(nb.RayBundle(**rays).to("cuda"), is_training=True, background_rgb=bg.cuda(), global_step=10)
    assert float(out.visibilities.abs().max()) == 0.0 and float(out.specular_cue.abs().max()) == 0.0
    assert torch.isfinite(out.rgb).all()


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("kind", ["init", "sharp"])
def test_sdf_query_matches_oracle(kind, impl):
    """SDFNetwork.sdf / .gradient / .forward on arbitrary points, ragged N (not a multiple of the tile)."""
    cfg = nb.NeuSModelConfig()
    sd = T.make_state(kind, cfg)
    m = nb.NeuSHintRenderer(cfg, mlp_impl=impl); m.load_state_dict(sd); m.cuda()
    g = torch.Generator().manual_seed(1)
    pts = torch.cat([(torch.rand(1000, 3, generator=g) - 0.5) * 2.6,                 # inside the unit sphere region
                     4.5 * torch.nn.functional.normalize(torch.randn(77, 3, generator=g), dim=-1)])   # at light distance
    ocfg = orc.OracleConfig.from_model_config(cfg)
    W = orc.effective_weights(sd)
    want = orc.sdf_mlp(W, pts, ocfg, want_feat=True, want_grad=True)
    full = m.sdf_network(pts.cuda())
    grad = m.sdf_network.gradient(pts.cuda())
    sdf = m.sdf_network.sdf(pts.cuda())
    assert full.shape == (1077, 257) and grad.shape == (1077, 1, 3) and sdf.shape == (1077, 1)
    # fp32 noise grows with |x| (Fourier arguments up to 430 rad at the light distance)
    k = 1.0 if impl == "fp32" else 4.0            # the split-fp16 tensor-core engine carries ~2e-5 absolute noise
    assert (sdf.cpu() - want["sdf"]).abs().max() < 2e-5 * k
    assert (full[:, 1:].cpu() - want["feat"]).abs().max() < 2e-4 * k
    assert (grad[:, 0].cpu() - want["grad"]).abs().max() < (2e-3 if kind == "sharp" else 3e-4)
    w64 = orc.sdf_mlp(orc.effective_weights(sd, torch.float64), pts.double(), ocfg, want_grad=True)
    print(kind, impl, "sdf err vs fp64:", float((sdf.cpu().double() - w64["sdf"]).abs().max()),
          "grad err vs fp64:", float((grad[:, 0].cpu().double() - w64["grad"]).abs().max()))


@torch.no_grad()
def test_edge_cases():
    """R = 1, R not a multiple of anything, rays that miss the sphere, R = 0."""
    cfg = nb.NeuSModelConfig()
    m = nb.NeuSHintRenderer(cfg, mlp_impl="auto"); m.load_state_dict(T.make_state("init", cfg)); m.cuda()
    rays = orc.synthetic_rays(37, seed=2, crop=800)
    full = m(nb.RayBundle(**rays).to("cuda"), background_rgb=torch.ones(1, 3).cuda())
    one = m(nb.RayBundle(**{k: v[5:6] for k, v in rays.items()}).to("cuda"), background_rgb=torch.ones(1, 3).cuda())
    assert torch.allclose(one.rgb, full.rgb[5:6], atol=1e-5)           # rays are independent
    empty = m(nb.RayBundle(**{k: v[:0] for k, v in rays.items()}).to("cuda"))
    assert empty.rgb.shape == (0, 3)
    assert torch.equal(full.relax_inside_sphere, full.inside_sphere)     # reference quirk Q1


@torch.no_grad()
def test_full_size_properties():
    """BASELINE.json config #2 size (4096 x 128): size-independent properties."""
    cfg = nb.NeuSModelConfig()
    m = nb.NeuSHintRenderer(cfg, mlp_impl="auto"); m.load_state_dict(T.make_state("init", cfg)); m.cuda()
    rays = orc.synthetic_rays(4096, seed=3407, crop=800)
    b = nb.RayBundle(**rays).to("cuda")
    out = m(b, background_rgb=torch.ones(1, 3).cuda(), return_extras=True)
    assert torch.isfinite(out.rgb).all() and out.rgb.min() >= 0 and out.rgb.max() <= 1.0 + 1e-5
    assert (out.weights >= 0).all() and (out.weights.sum(-1) <= 1.0 + 1e-4).all()
    assert (out.z_vals[:, 1:] >= out.z_vals[:, :-1]).all()                               # sortedness
    assert (out.visibilities >= 0).all() and (out.visibilities <= 1.0 + 1e-5).all()
    n = out.normalized_analytic_normals.norm(dim=-1)
    assert (n - 1).abs().max() < 1e-4
    # determinism + batch-composition independence (a chunked render equals the full one)
    out2 = m(b, background_rgb=torch.ones(1, 3).cuda())
    assert torch.equal(out.rgb, out2.rgb)
    chunk = m(nb.RayBundle(**{k: v[512:1024] for k, v in rays.items()}).to("cuda"), background_rgb=torch.ones(1, 3).cuda())
    assert torch.allclose(chunk.rgb, out.rgb[512:1024], atol=1e-5)
    # a 512-ray slice against the oracle
    sl = {k: v[:256] for k, v in rays.items()}
    ocfg = orc.OracleConfig.from_model_config(cfg)
    with torch.no_grad():
        want = orc.render_forward(T.make_state("init", cfg), ocfg, sl["origins"], sl["directions"], sl["pl_positions"],
                                  sl["nears"], sl["fars"], background_rgb=torch.ones(1, 3))
    assert (out.rgb[:256].cpu() - want["rgb"]).abs().max() < 1e-3
    assert (out.depth[:256].cpu() - want["depth"]).abs().max() < 1e-3


# Limits per case / engine / tensor group as [max-abs error, L2 error], both relative to the tensor (max |g|, ||g||).  Measured on
# B200 (gpurun_out/r2_07_pytest_grad.log), worst tensor of each group:
#   init weights, cos_anneal 0.5 (16 rays): fp32 engine 4.7e-4 / 1.9e-4; tcgen05 engine 1.4e-3 / 7.4e-4; tensors behind the fp16
#     reflectance backward (color_network.*, sdf_network.out_feat.*, light positions) 2.1e-2 / 7.8e-3
#   sharp weights, step 60000 (24 rays, inv_s ~ 403): fp32 engine 3.6e-3 / 2.5e-3 and 6.6e-3 for the variance -- that is the fp32
#     noise floor of the case (the oracle itself sits 2-3e-3 from the reference fixture: a displaced importance sample changes
#     which points a ray evaluates); tcgen05 engine 7.7e-3 / 7.2e-3, reflectance group 1.4e-2 / 5.7e-3
GRAD_TOL = {
    "init": {"fp32": dict(max=1.5e-3, l2=1e-3), "auto": dict(max=4e-3, l2=2.5e-3), "auto_color": dict(max=4e-2, l2=1.5e-2)},
    "sharp": {"fp32": dict(max=1.2e-2, l2=1.2e-2), "auto": dict(max=1.5e-2, l2=1.5e-2), "auto_color": dict(max=3e-2, l2=1.5e-2)},
}
VERBOSE_GRADS = False


def _grad_group(pname, impl):
    """Tensors whose gradient passes through the reflectance MLP's fp16 backward (tcgen05 engine) carry its noise."""
    impl = "auto" if impl == "auto-composed" else impl
    if impl != "fp32" and (pname.startswith("color_network.") or pname.startswith("sdf_network.out_feat") or pname == "ray::pl_positions"):
        return "auto_color"
    return impl




@torch.no_grad()
def test_full_size_sharp_all_fields_match_oracle():
    """BASELINE.json config #2 size (4096 rays x 128 samples) with the trained-like sharp weights (inv_s ~ 403, perturbed
    non-spherical field): EVERY RenderOutput field of the CUDA path against the oracle (evaluated in 512-ray chunks, the
    reference's inference_chunk_size), at the same tolerances as the small sharp fixture."""
    cfg = nb.NeuSModelConfig()
    sd = T.make_state("sharp", cfg)
    m = nb.NeuSHintRenderer(cfg, mlp_impl="auto"); m.load_state_dict(sd); m.cuda()
    rays = orc.synthetic_rays(4096, seed=4242, crop=800)
    out = m(nb.RayBundle(**rays).to("cuda"), background_rgb=torch.ones(1, 3).cuda(), return_extras=True)
    got = T.to_np(out)
    ocfg = orc.OracleConfig.from_model_config(cfg)
    parts = []
    for i0 in range(0, 4096, 512):
        sl = {k: v[i0:i0 + 512] for k, v in rays.items()}
        parts.append(T.to_np(orc.render_forward(sd, ocfg, sl["origins"], sl["directions"], sl["pl_positions"], sl["nears"], sl["fars"],
                                                background_rgb=torch.ones(1, 3))))
    want = {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}
    stats = T.compare_outputs(got, want, label="cuda[auto]-vs-oracle[4096 sharp]", **T.TOL_TC["sharp"])
    assert stats["psnr_between"] > 60
    print("FULLSHARP", {k: f"{v:.2e}" for k, v in stats.items()})


@pytest.mark.parametrize("impl", IMPLS + ["auto-composed"])
@pytest.mark.parametrize("name", list(T.GRAD_CASES) + ["outside_train_64x128"])
def test_training_gradients_match_oracle(name, impl):
    """Training-mode forward + loss.backward() through the CUDA path (the fused node by default, the composed autograd nodes as
    "auto-composed", and -- with the outside NeRF -- the composed route plus the torch evaluation of the background model: all 70
    parameter tensors incl. outside_nerf.*) against the oracle's autograd gradients -- which tests/test_oracle_golden.py pins to the
    reference's own loss.backward() fixtures -- for all 46 parameter tensors INCLUDING deviation_network.variance and for the ray
    inputs (origins / directions / light positions).  Two cases: init weights half way through annealing, and trained-like sharp
    weights (inv_s ~ 403) at global_step 60000 (cos_anneal = 1)."""
    # outside NeRF: the training-mode fixture case with 64 rays instead of 8 (at 8 rays the parameter gradients are ~1e-6 and the
    # tcgen05 engine's fp16x3 noise at inv_s ~ 403 reaches 2-6 % of them; measured at 64 rays: <= 9e-3, gpurun_out r2 logs)
    case = dict(T.CASES["outside_train_8x128"], R=64) if name == "outside_train_64x128" else T.CASES[name]
    m, cfg, sd = build_module(case, impl)
    rays, bg = T.case_inputs(case)
    R = case["R"]
    if cfg.renderer.use_outside_nerf and impl == "auto-composed":
        pytest.skip("with the outside NeRF the tcgen05 engine always takes the composed route (covered by impl 'auto')")
    torch.manual_seed(7)
    jp = torch.rand([R, 1], device="cuda")
    jo = torch.rand([R, cfg.renderer.n_outside_samples], device="cuda") if cfg.renderer.use_outside_nerf else None
    js = torch.rand([R, cfg.renderer.n_shadow_samples], device="cuda")
    torch.manual_seed(7)
    dev_rays = {k: v.cuda() for k, v in rays.items()}
    for k in ("origins", "directions", "pl_positions"):
        dev_rays[k].requires_grad_(True)
    out = m(nb.RayBundle(**dev_rays), is_training=True, background_rgb=bg.cuda(), global_step=case["global_step"])
    assert out.rgb.requires_grad and out.weights.requires_grad and out.analytic_normals.requires_grad and out.s_val.requires_grad
    assert not out.depth.requires_grad and not out.visibilities.requires_grad
    gt = torch.rand(R, 3, generator=torch.Generator().manual_seed(T.GRAD_CASES.get(name, 79)))
    loss = orc.training_loss({"rgb": out.rgb, "analytic_normals": out.analytic_normals,
                              "relax_inside_sphere": out.relax_inside_sphere}, gt.cuda())
    loss.backward()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    leaves = {k: rays[k].clone().requires_grad_(True) for k in ("origins", "directions", "pl_positions")}
    ocfg = orc.OracleConfig.from_model_config(cfg)
    want = orc.render_forward(sdr, ocfg, leaves["origins"], leaves["directions"], leaves["pl_positions"], rays["nears"], rays["fars"],
                              is_training=True, background_rgb=bg, cos_anneal=min(1.0, case["global_step"] / cfg.anneal_end),
                              jitter_primary=jp.cpu(), jitter_shadow=js.cpu(), jitter_outside=jo.cpu() if jo is not None else None)
    loss_o = orc.training_loss(want, gt)
    assert abs(float(loss) - float(loss_o)) < 1e-4
    keys = sorted(sdr)
    allg = torch.autograd.grad(loss_o, [sdr[k] for k in keys] + list(leaves.values()))
    go = dict(zip(keys, allg[:len(keys)]))
    go.update({"ray::" + k: g for k, g in zip(leaves, allg[len(keys):])})
    got = {n_: p.grad for n_, p in m.named_parameters()}
    got.update({"ray::" + k: dev_rays[k].grad for k in leaves})
    worst, failures = {}, []
    for pname, g_cuda in got.items():
        assert g_cuda is not None and torch.isfinite(g_cuda).all(), pname          # DDP find_unused_parameters=False is safe
        g = go[pname]
        d = g_cuda.detach().cpu() - g
        e_max = float(d.abs().max() / g.abs().max().clamp_min(1e-8))
        e_l2 = float(d.norm() / g.norm().clamp_min(1e-8))
        grp = _grad_group(pname, impl)
        w = worst.setdefault(grp, [0.0, 0.0, ""])
        if e_max > w[0]:
            w[0], w[2] = e_max, pname
        w[1] = max(w[1], e_l2)
        tol = GRAD_TOL[case["weights"]][grp]
        if not (e_max < tol["max"] and e_l2 < tol["l2"]):
            failures.append(f"{pname}: max-abs {e_max:.2e} (limit {tol['max']:.0e}), L2 {e_l2:.2e} (limit {tol['l2']:.0e})")
        if VERBOSE_GRADS:
            print(f"GRAD {name} {impl} {pname:45s} max {e_max:.2e} l2 {e_l2:.2e}")
    print("GRADERR", name, impl, {k: (f"{v[0]:.2e}", f"{v[1]:.2e}", v[2]) for k, v in worst.items()})
    assert not failures, f"{name}[{impl}]: " + "; ".join(failures)
    assert "deviation_network.variance" in got and float(go["deviation_network.variance"].abs()) > 0


@pytest.mark.parametrize("kind", ["init", "sharp"])
def test_camera_gradients_match_oracle(kind):
    """cam-opt / register_view (pipelines/base_pipeline.py:71-91): with the network frozen, the gradients of an image loss
    w.r.t. origins, directions and light positions equal the oracle's autograd gradients (values, not just non-zero)."""
    case = dict(T.CASES["cfg1_64x32"], weights=kind)
    m, cfg, sd = build_module(case, "auto")
    for p in m.parameters():
        p.requires_grad_(False)
    rays, bg = T.case_inputs(case)
    dev = {k: v.cuda().requires_grad_(True) for k, v in rays.items()}
    gt = torch.rand(case["R"], 3, generator=torch.Generator().manual_seed(5))
    out = m(nb.RayBundle(**dev), is_training=False, background_rgb=bg.cuda())
    (out.rgb - gt.cuda()).abs().sum().backward()
    leaves = {k: rays[k].clone().requires_grad_(True) for k in ("origins", "directions", "pl_positions")}
    want = orc.render_forward(sd, orc.OracleConfig.from_model_config(cfg), leaves["origins"], leaves["directions"], leaves["pl_positions"],
                              rays["nears"], rays["fars"], background_rgb=bg)
    go = torch.autograd.grad((want["rgb"] - gt).abs().sum(), list(leaves.values()))
    # measured on B200: init 1.0e-4 (origins) / 9.4e-5 (directions) / 2.1e-2 (light positions: their gradient runs through the
    # fp16 reflectance backward only); sharp <= 9.3e-3 / ... (fp32 noise floor of the sharp field, see GRAD_TOL)
    lim = {"init": dict(origins=2e-3, directions=2e-3, pl_positions=4e-2), "sharp": dict(origins=2e-2, directions=2e-2, pl_positions=4e-2)}[kind]
    errs = {}
    for k, g in zip(leaves, go):
        got = dev[k].grad
        assert got is not None and torch.isfinite(got).all() and float(g.abs().max()) > 0, k
        errs[k] = float((got.cpu() - g).abs().max() / g.abs().max())
    print("CAMGRAD", kind, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, e in errs.items():
        assert e < lim[k], f"{kind} d loss / d {k}: rel err {e:.2e} (limit {lim[k]:.0e})"
    # the reference's final z_vals are produced under no_grad (models/neus_hint_model.py:696-713): near / far get no gradient
    assert dev["nears"].grad is None or float(dev["nears"].grad.abs().max()) == 0.0


@torch.no_grad()
def test_render_maps_equal_host_reduction():
    """render_maps (device-side reduction) == the einsum the reference pipeline applies to the per-sample tensors."""
    case = T.CASES["cfg2_32x128"]
    m, cfg, sd = build_module(case, "auto")
    rays, bg = T.case_inputs(case)
    b = nb.RayBundle(**rays).to("cuda")
    full = m(b, background_rgb=bg.cuda())
    maps = m.render_maps(b, background_rgb=bg.cuda())
    assert torch.equal(maps["rgb"], full.rgb) and torch.equal(maps["depth"], full.depth)
    assert torch.equal(maps["shadow_map"], full.visibilities)
    assert torch.allclose(maps["specular_hint"], full.specular_cue[:, 0, :])
    want = torch.einsum("...ij,...i,...i->...j", full.analytic_normals, full.weights, full.inside_sphere)
    wantn = torch.einsum("...ij,...i,...i->...j", full.normalized_analytic_normals, full.weights, full.inside_sphere)
    assert torch.allclose(maps["analytic_normals"], want, atol=2e-6)
    assert torch.allclose(maps["normalized_analytic_normals"], wantn, atol=2e-6)
    assert m.last_launch_count < 29


@torch.no_grad()
def test_render_to_host_equals_forward():
    """render_to_host (geometry block copied on a second stream behind NrhOutputs.early_event) delivers exactly what
    forward() + .cpu() does (the reference's evaluation hand-off, pipelines/base_pipeline.py:114-120)."""
    m, cfg, sd = build_module(T.CASES["cfg2_32x128"], "auto")
    from nrhints_b200.workload import synthetic_rays
    rays = nb.RayBundle(**synthetic_rays(1024, seed=11)).pin_memory()
    bg = torch.ones(1, 3)
    want = m(rays.to("cuda"), background_rgb=bg.cuda())
    for _ in range(3):
        got = m.render_to_host(rays, background_rgb=bg)
    for k, v in want.as_dict().items():
        g = getattr(got, k)
        if v is None:
            assert g is None
            continue
        assert g.device.type == "cpu" and g.is_pinned(), k
        assert torch.equal(g, v.cpu()), k


@torch.no_grad()
def test_render_image_equals_chunked_render_maps():
    """render_image (whole view in one call, ragged last slice, host rays staged on a second stream) delivers exactly the
    per-ray maps of separate render_maps calls -- the reference's per-chunk evaluation loop, pipelines/base_pipeline.py:107-133."""
    m, cfg, sd = build_module(T.CASES["cfg2_32x128"], "auto")
    from nrhints_b200.workload import synthetic_rays
    N = 1000
    rays = nb.RayBundle(**synthetic_rays(N, seed=5)).pin_memory()
    bg = torch.ones(1, 3)
    want = m.render_maps(rays.to("cuda"), background_rgb=bg.cuda())
    for src in (rays, rays.to("cuda")):
        got = m.render_image(src, background_rgb=bg, chunk_rays=384)
        assert set(got) == set(want)
        for k, v in want.items():
            assert not got[k].is_cuda and got[k].shape == v.shape
            assert torch.allclose(got[k], v.cpu(), atol=1e-5), k         # chunk composition changes nothing (rays are independent)
    dev = m.render_image(rays, background_rgb=bg, chunk_rays=384, to_host=False)
    assert dev["rgb"].is_cuda and torch.allclose(dev["rgb"].cpu(), want["rgb"].cpu(), atol=1e-5)


@torch.no_grad()
def test_host_results_of_consecutive_calls_do_not_alias():
    """The reference's evaluation loop keeps `rendering_res.to('cpu')` of every 512-ray chunk in a list and concatenates at the
    end (pipelines/base_pipeline.py:112-123): results handed out by render_to_host / to_host / render_image must stay valid
    after later calls with the same shapes."""
    m, cfg, sd = build_module(T.CASES["cfg2_32x128"], "auto")
    from nrhints_b200.workload import synthetic_rays
    bg = torch.ones(1, 3)
    chunks = [nb.RayBundle(**synthetic_rays(256, seed=s)).pin_memory() for s in (1, 2, 3)]
    want = [m(c.to("cuda"), background_rgb=bg.cuda()) for c in chunks]
    got = [m.render_to_host(c, background_rgb=bg) for c in chunks]           # same shapes every call
    got2 = [m.to_host(w) for w in want]
    for w, g, g2 in zip(want, got, got2):
        for k in ("rgb", "weights", "analytic_normals", "visibilities", "specular_cue"):
            assert torch.equal(getattr(g, k), getattr(w, k).cpu()), k
            assert torch.equal(getattr(g2, k), getattr(w, k).cpu()), k
    assert got[0].rgb.data_ptr() != got[1].rgb.data_ptr() and got[0].weights.data_ptr() != got[2].weights.data_ptr()
    imgs = [m.render_image(c, background_rgb=bg) for c in chunks]
    maps = [m.render_maps(c.to("cuda"), background_rgb=bg.cuda()) for c in chunks]
    for im, mp in zip(imgs, maps):
        assert torch.allclose(im["rgb"], mp["rgb"].cpu(), atol=1e-6) and torch.allclose(im["depth"], mp["depth"].cpu(), atol=1e-6)
