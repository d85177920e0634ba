"""Developer sweep (GPU box, developer build): fp16 MMA passes per product outside the forward SDF layers.

    python tests/dev/rev_passes_sweep.py [quick]

For every setting of SdfTcParams::rev_passes / feat_passes (dev codes 10..14) it prints the parity statistics of the small fixtures
cases and of the 4096-ray sharp case against the oracle (same gates as tests/test_gpu_parity.py, failures reported, not raised) and
the device time of the fine pass (sdf + grad + feat), the shadow fine pass (sdf + grad) and the whole 4096 x 128 forward; then the
same for SdfBwdParams::passes (dev codes 20..22) on the fused SDF backward (gradient error against torch.autograd + time)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import tc_dev                                                   # noqa: E402  (selects the developer build)
import numpy as np                                              # noqa: E402
import torch                                                    # noqa: E402
import nrh_testlib as T                                         # noqa: E402
import nrhints_b200 as nb                                       # noqa: E402
from nrhints_b200 import autograd_fine, sdf_autograd            # noqa: E402
from oracle import nrh_oracle as orc                            # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
NAMES = {0: "rev3 feat3", 14: "rev3 feat1", 10: "rev2 feat3", 12: "rev2 feat1", 11: "rev1 feat3", 13: "rev1 feat1"}


def ev_time(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def fmt(d):
    return {k: f"{v:.1e}" for k, v in d.items()}


# ---------------- oracle side, once ----------------
small = ["cfg2_32x128", "sharp_32x128", "shadowonly_16x64"]
want_small = {n: T.to_np(T.run_oracle(T.CASES[n])) for n in small}
cfg = nb.NeuSModelConfig()
sd_sharp = T.make_state("sharp", cfg)
ocfg = orc.OracleConfig.from_model_config(cfg)
rays_big = orc.synthetic_rays(4096, seed=4242, crop=800)
want_big = None
if not quick:
    parts = []
    with torch.no_grad():
        for i0 in range(0, 4096, 512):
            sl = {k: v[i0:i0 + 512] for k, v in rays_big.items()}
            parts.append(T.to_np(orc.render_forward(sd_sharp, ocfg, sl["origins"], sl["directions"], sl["pl_positions"], sl["nears"], sl["fars"],
                                                    background_rgb=torch.ones(1, 3))))
    want_big = {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}
print("oracle done", flush=True)

m_big = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m_big.load_state_dict(sd_sharp); m_big.cuda()
bundle_big = nb.RayBundle(**rays_big).to("cuda")
bg1 = torch.ones(1, 3).cuda()
pts = (torch.rand(4096 * 128, 3, device="cuda") - 0.5) * 2

with torch.no_grad():
    for dbg, label in NAMES.items():
        tc_dev.configure(gen=1, dbg=dbg)
        print(f"==== {label} (dev code {dbg})", flush=True)
        for n in small:
            case = T.CASES[n]
            ccfg = T.make_config(case)
            m = nb.NeuSHintRenderer(ccfg, mlp_impl="tcgen05"); m.load_state_dict(T.make_state(case["weights"], ccfg)); m.cuda()
            rays, bg = T.case_inputs(case)
            out = m(nb.RayBundle(**rays).to("cuda"), background_rgb=bg.cuda(), return_extras=True)
            try:
                st = T.compare_outputs(T.to_np(out), want_small[n], label=n, **T.TOL_TC[case["weights"]])
                print("  ", n, "ok", fmt(st), flush=True)
            except AssertionError as e:
                print("  ", n, "GATE FAIL:", e, flush=True)
        if want_big is not None:
            out = m_big(bundle_big, background_rgb=bg1, return_extras=True)
            got = T.to_np(out)
            try:
                st = T.compare_outputs(got, want_big, label="4096 sharp", **T.TOL_TC["sharp"])
                print("   4096-sharp ok", fmt(st), flush=True)
            except AssertionError as e:
                print("   4096-sharp GATE FAIL:", e, flush=True)
                try:
                    st = T.compare_outputs(got, want_big, label="4096 sharp (loose)", per_ray_tol=1, per_sample_tol=1, normals_tol=1, max_displaced_frac=1.0)
                    print("   4096-sharp raw stats", fmt(st), flush=True)
                except AssertionError as e2:
                    print("   ", e2)
        t_fine = ev_time(lambda: m_big.sdf_query(pts, want_grad=True, want_feat=True))
        t_shadow = ev_time(lambda: m_big.sdf_query(pts, want_grad=True))
        t_fwd = ev_time(lambda: m_big(bundle_big, background_rgb=bg1))
        print(f"   time: fine (sdf+grad+feat) {t_fine:.3f} ms, sdf+grad {t_shadow:.3f} ms, forward 4096x128 {t_fwd:.3f} ms = {4096 / t_fwd:.1f} K rays/s", flush=True)

# ---------------- training backward ----------------
tc_dev.configure(gen=1, dbg=0)


def reference(m, p, d_sdf, d_feat, d_grad):
    w = m._autograd_weights()
    head = {"sdf_w": w["sdf_w_head"], "sdf_b": w["sdf_b_head"], "feat_w": w["feat_w"], "feat_b": w["feat_b"]}
    x = p.clone().requires_grad_(True)
    out = autograd_fine.sdf_forward(w["sdf_w"], w["sdf_b"], head, x)
    sdf, feat = out[:, :1], out[:, 1:]
    grad = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=True)[0]
    loss = (sdf * d_sdf).sum() + (feat * d_feat).sum() + (grad * d_grad).sum()
    params = [q for n, q in m.named_parameters() if n.startswith("sdf_network.")]
    gs = torch.autograd.grad(loss, [x] + params)
    return gs[0], gs[1:]


for kind, N in (("init", 1000), ("sharp", 384)):
    m = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m.load_state_dict(T.make_state(kind, cfg)); m.cuda()
    g = torch.Generator().manual_seed(N)
    p = ((torch.rand(N, 3, generator=g) - 0.5) * 2.0).cuda()
    d_sdf = (torch.randn(N, 1, generator=g) * 0.3).cuda()
    d_feat = (torch.randn(N, 256, generator=g) * 0.01).cuda()
    d_grad = (torch.randn(N, 3, generator=g) * 0.1).cuda()
    dpts_r, gp_r = reference(m, p, d_sdf, d_feat, d_grad)
    params = [q for n, q in m.named_parameters() if n.startswith("sdf_network.")]
    for dbg, label in ((22, "bwd3"), (20, "bwd2"), (21, "bwd1")):
        tc_dev.configure(gen=1, dbg=dbg)
        x = p.clone().requires_grad_(True)
        sdf, feat, grad = sdf_autograd.sdf_fine(m, x, m._autograd_weights())
        loss = (sdf * d_sdf).sum() + (feat * d_feat).sum() + (grad * d_grad).sum()
        gs = torch.autograd.grad(loss, [x] + params)
        e_pts = float((gs[0] - dpts_r).abs().max() / dpts_r.abs().max())
        worst = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-12)) for a, b in zip(gs[1:], gp_r))
        print(f"backward {kind} N={N} {label}: d_pts rel err {e_pts:.2e}, worst parameter rel err {worst:.2e}", flush=True)

m = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m.load_state_dict(sd_sharp); m.cuda()
params = [q for n, q in m.named_parameters() if n.startswith("sdf_network.")]
Nb = 4096 * 128
pb = ((torch.rand(Nb, 3, device="cuda") - 0.5) * 2.0)
for dbg, label in ((22, "bwd3"), (20, "bwd2"), (21, "bwd1")):
    tc_dev.configure(gen=1, dbg=dbg)

    def step():
        x = pb.clone().requires_grad_(True)
        sdf, feat, grad = sdf_autograd.sdf_fine(m, x, m._autograd_weights())
        loss = sdf.sum() + feat.sum() * 0.01 + grad.sum() * 0.1
        torch.autograd.grad(loss, [x] + params)
    t = ev_time(step, n=5, warm=2)
    print(f"fused SDF forward-with-tape + backward + weight gradients over {Nb} points, {label}: {t:.3f} ms", flush=True)
