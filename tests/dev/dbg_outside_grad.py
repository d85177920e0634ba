import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import nrh_testlib as T
import nrhints_b200 as nb
from oracle import nrh_oracle as orc
base = T.CASES["outside_train_8x128"]
for R in (8, 64):
    case = dict(base, R=R)
    cfg = T.make_config(case)
    sd = T.make_state(case["weights"], cfg)
    rays, bg = T.case_inputs(case)
    gt = torch.rand(R, 3, generator=torch.Generator().manual_seed(79))
    res = {}
    for impl in ("fp32", "auto"):
        m = nb.NeuSHintRenderer(cfg, mlp_impl=impl); m.load_state_dict(sd); m.cuda()
        torch.manual_seed(7)
        jp = torch.rand([R, 1], device="cuda"); jo = torch.rand([R, cfg.renderer.n_outside_samples], device="cuda"); js = torch.rand([R, cfg.renderer.n_shadow_samples], device="cuda")
        torch.manual_seed(7)
        dev = {k: v.cuda() for k, v in rays.items()}
        out = m(nb.RayBundle(**dev), is_training=True, background_rgb=bg.cuda(), global_step=case["global_step"])
        loss = orc.training_loss({"rgb": out.rgb, "analytic_normals": out.analytic_normals, "relax_inside_sphere": out.relax_inside_sphere}, gt.cuda())
        loss.backward()
        res[impl] = (float(loss), {n: p.grad.detach().cpu() for n, p in m.named_parameters()}, out.weights.detach().cpu(), out.rgb.detach().cpu())
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ocfg = orc.OracleConfig.from_model_config(cfg)
    want = orc.render_forward(sdr, ocfg, rays["origins"], rays["directions"], rays["pl_positions"], rays["nears"], rays["fars"], is_training=True,
                              background_rgb=bg, cos_anneal=1.0, jitter_primary=jp.cpu(), jitter_shadow=js.cpu(), jitter_outside=jo.cpu())
    lo = orc.training_loss(want, gt)
    keys = sorted(sdr)
    go = dict(zip(keys, torch.autograd.grad(lo, [sdr[k] for k in keys])))
    print("R", R, "loss fp32/auto/oracle", res["fp32"][0], res["auto"][0], float(lo))
    print("  weights: max |auto - fp32|", float((res["auto"][2] - res["fp32"][2]).abs().max()), "max |fp32 - oracle|", float((res["fp32"][2] - want["weights"].detach()).abs().max()),
          " rgb: max |auto - fp32|", float((res["auto"][3] - res["fp32"][3]).abs().max()))
    for n in ("deviation_network.variance", "sdf_network.lin0.weight_v", "sdf_network.lin4.weight_v", "sdf_network.lin7.weight_v", "sdf_network.out_sdf.weight_v",
              "color_network.lin0.weight_v", "outside_nerf.pts_linears.0.weight", "outside_nerf.rgb_linear.weight"):
        g = go[n]
        e = lambda a: float((a - g).abs().max() / g.abs().max().clamp_min(1e-12))
        print(f"  {n:40s} fp32-vs-oracle {e(res['fp32'][1][n]):.2e}  auto-vs-oracle {e(res['auto'][1][n]):.2e}  auto-vs-fp32 {float((res['auto'][1][n]-res['fp32'][1][n]).abs().max()/g.abs().max().clamp_min(1e-12)):.2e}  |g|max {float(g.abs().max()):.2e}")
