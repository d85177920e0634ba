import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import nrh_testlib as T
import nrhints_b200 as nb
from oracle import nrh_oracle as orc
for impl in ("auto", "fp32"):
    for kind in ("init", "sharp"):
        cfg = nb.NeuSModelConfig()
        m = nb.NeuSHintRenderer(cfg, mlp_impl=impl); m.load_state_dict(T.make_state(kind, cfg)); m.cuda()
        rays = nb.RayBundle(**orc.synthetic_rays(256, seed=5)).to("cuda")
        outs = []
        with torch.no_grad():
            for rep in range(4):
                if rep == 2:
                    m._packed_key = None
                o = m(rays, background_rgb=torch.ones(1, 3).cuda(), return_extras=True)
                outs.append({k: v.clone() for k, v in o.as_dict().items() if v is not None})
        for rep in range(1, 4):
            diffs = {k: float((outs[rep][k] - outs[0][k]).abs().max()) for k in outs[0]}
            print(impl, kind, "rep", rep, {k: f"{v:.1e}" for k, v in diffs.items() if v > 0})
