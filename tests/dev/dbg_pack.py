import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import nrh_testlib as T
import nrhints_b200 as nb
from nrhints_b200 import fused_step
for kind in ("init", "sharp"):
    cfg = nb.NeuSModelConfig()
    m = nb.NeuSHintRenderer(cfg, mlp_impl="auto"); m.load_state_dict(T.make_state(kind, cfg)); m.cuda()
    dev = torch.device("cuda")
    from oracle import nrh_oracle as orc
    rays = nb.RayBundle(**orc.synthetic_rays(256, seed=5)).to("cuda")
    with torch.no_grad():
        oa = m(rays, background_rgb=torch.ones(1, 3).cuda()).rgb.clone()
    a = m._ensure_packed(dev).clone()
    m._packed_key = None
    orig = fused_step.can_pack_wn
    fused_step.can_pack_wn = lambda r: False
    b = m._ensure_packed(dev).clone()
    with torch.no_grad():
        ob = m(rays, background_rgb=torch.ones(1, 3).cuda()).rgb.clone()
    print("rgb diff", float((oa - ob).abs().max()))
    fused_step.can_pack_wn = orig
    n = min(a.numel(), b.numel()) // 4
    fa, fb = a[:n * 4].view(torch.float32), b[:n * 4].view(torch.float32)
    ok = torch.isfinite(fa) & torch.isfinite(fb)
    d = (fa - fb).abs()
    d[~ok] = 0
    rel = d / fb.abs().clamp_min(1e-6)
    rel[~ok] = 0
    i = int(rel.argmax())
    print(kind, "bytes", a.numel(), b.numel(), "max abs", float(d.max()), "max rel", float(rel.max()), "at float", i, float(fa[i]), float(fb[i]),
          "n differing", int((d > 0).sum()), "of", n, "bitwise equal bytes", float((a == b).float().mean()))
