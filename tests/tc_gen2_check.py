"""Developer check of the second-generation tcgen05 engine (mlp_tc2.inc) against generation 1 and the fp64 oracle; GPU box."""
import os, sys, time
import torch
torch.set_grad_enabled(False)
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import nrh_testlib as T
import nrhints_b200 as nb
from oracle import nrh_oracle as orc

kind = sys.argv[1] if len(sys.argv) > 1 else "init"
cfg = nb.NeuSModelConfig()
sd = T.make_state(kind, cfg)
m = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m.load_state_dict(sd); m.cuda()
g = torch.Generator().manual_seed(1)
pts = torch.cat([(torch.rand(1000, 3, generator=g) - 0.5) * 2.6, 4.5 * torch.nn.functional.normalize(torch.randn(77, 3, generator=g), dim=-1)])
ocfg = orc.OracleConfig.from_model_config(cfg)
w64 = orc.sdf_mlp(orc.effective_weights(sd, torch.float64), pts.double(), ocfg, want_feat=True, want_grad=True)
modes = [(False, False), (True, False), (True, True), (False, True)]
for gen in ("1", "2"):
    os.environ["NRH_TC_GEN"] = gen
    for wg, wf in modes:
        sdf, grad, feat = m.sdf_query(pts.cuda(), want_grad=wg, want_feat=wf)
        torch.cuda.synchronize()
        e = ["sdf %.2e" % float((sdf.cpu().double() - w64["sdf"][:, 0]).abs().max())]
        if grad is not None: e.append("grad %.2e" % float((grad.cpu().double() - w64["grad"]).abs().max()))
        if feat is not None: e.append("feat %.2e" % float((feat.cpu().double() - w64["feat"]).abs().max()))
        print(f"gen {gen} grad={wg} feat={wf}: err vs fp64", e, flush=True)
big = (torch.rand(4096 * 128, 3, device="cuda") - 0.5) * 2
for gen in ("1", "2"):
    os.environ["NRH_TC_GEN"] = gen
    for wg, wf in modes:
        for _ in range(2): m.sdf_query(big, want_grad=wg, want_feat=wf)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): m.sdf_query(big, want_grad=wg, want_feat=wf)
        b.record(); torch.cuda.synchronize()
        print(f"gen {gen} grad={wg} feat={wf}: {a.elapsed_time(b) / 5:.3f} ms / 524288 points", flush=True)
rays = nb.RayBundle(**orc.synthetic_rays(4096, seed=3407)).to("cuda")
bg = torch.ones(1, 3).cuda()
outs = {}
for gen in ("1", "2"):
    os.environ["NRH_TC_GEN"] = gen
    for _ in range(2): outs[gen] = m(rays, background_rgb=bg)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): m(rays, background_rgb=bg)
    b.record(); torch.cuda.synchronize()
    print(f"gen {gen} full forward: {a.elapsed_time(b) / 5:.3f} ms / 4096 rays", flush=True)
print("rgb gen2 vs gen1 max abs diff", float((outs["1"].rgb - outs["2"].rgb).abs().max()))
