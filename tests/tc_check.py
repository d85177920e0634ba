"""Developer check of the tcgen05 engine against the fp32 engine and the oracle (run on the GPU box)."""
import sys, time
import torch
torch.set_grad_enabled(False)
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import nrh_testlib as T
import nrhints_b200 as nb
from oracle import nrh_oracle as orc

kind = sys.argv[1] if len(sys.argv) > 1 else "init"
cfg = nb.NeuSModelConfig()
sd = T.make_state(kind, cfg)
mods = {}
for impl in ("fp32", "tcgen05"):
    m = nb.NeuSHintRenderer(cfg, mlp_impl=impl); m.load_state_dict(sd); m.cuda(); mods[impl] = m
g = torch.Generator().manual_seed(1)
pts = torch.cat([(torch.rand(1000, 3, generator=g) - 0.5) * 2.6, 4.5 * torch.nn.functional.normalize(torch.randn(77, 3, generator=g), dim=-1)])
ocfg = orc.OracleConfig.from_model_config(cfg)
w64 = orc.sdf_mlp(orc.effective_weights(sd, torch.float64), pts.double(), ocfg, want_feat=True, want_grad=True)
for mode in [(False, False), (False, True), (True, False), (True, True)]:
    res = {}
    for impl, m in mods.items():
        sdf, grad, feat = m.sdf_query(pts.cuda(), want_grad=mode[0], want_feat=mode[1])
        torch.cuda.synchronize()
        res[impl] = (sdf, grad, feat)
        e = [float((sdf.cpu().double() - w64["sdf"][:, 0]).abs().max())]
        if grad is not None: e.append(float((grad.cpu().double() - w64["grad"]).abs().max()))
        if feat is not None: e.append(float((feat.cpu().double() - w64["feat"]).abs().max()))
        print(f"mode grad={mode[0]} feat={mode[1]} {impl:8s} err vs fp64 (sdf[,grad][,feat]):", ["%.2e" % x for x in e], flush=True)
rays = orc.synthetic_rays(256, seed=3, crop=300)
b = nb.RayBundle(**rays).to("cuda")
outs = {}
for impl, m in mods.items():
    outs[impl] = m(b, background_rgb=torch.ones(1, 3).cuda(), return_extras=True)
    torch.cuda.synchronize()
a, c = T.to_np(outs["fp32"]), T.to_np(outs["tcgen05"])
try:
    print("full forward tc vs fp32:", {k: "%.2e" % v for k, v in T.compare_outputs(c, a, label="tc-vs-fp32", max_displaced_frac=1.0, **T.TOL[kind]).items()})
except AssertionError as e:
    print("COMPARE FAIL", e)
want = T.to_np(orc.render_forward(sd, ocfg, rays["origins"], rays["directions"], rays["pl_positions"], rays["nears"], rays["fars"], background_rgb=torch.ones(1, 3), dtype=torch.float64))
for nm, o in (("fp32", a), ("tcgen05", c)):
    try:
        print(nm, "vs fp64 oracle:", {k: "%.2e" % v for k, v in T.compare_outputs(o, want, label=nm, max_displaced_frac=1.0, per_ray_tol=1, per_sample_tol=1, normals_tol=1).items()})
    except AssertionError as e:
        print("COMPARE FAIL", e)
print("sampled_color diff", float((outs["fp32"].sampled_color - outs["tcgen05"].sampled_color).abs().max()))
big = nb.RayBundle(**orc.synthetic_rays(4096, seed=3407)).to("cuda")
for impl, m in mods.items():
    for _ in range(2): m(big, background_rgb=torch.ones(1, 3).cuda())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): m(big, background_rgb=torch.ones(1, 3).cuda())
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"{impl}: {dt*1e3:.2f} ms / 4096 rays -> {4096/dt:.0f} rays/s")
