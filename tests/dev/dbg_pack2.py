import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import nrh_testlib as T
import nrhints_b200 as nb
from nrhints_b200 import fused_step
from oracle import nrh_oracle as orc
dev = torch.device("cuda")
cfg = nb.NeuSModelConfig()
m = nb.NeuSHintRenderer(cfg, mlp_impl="auto"); m.load_state_dict(T.make_state("sharp", cfg)); m.cuda()
rays = nb.RayBundle(**orc.synthetic_rays(256, seed=5)).to("cuda")
bg = torch.ones(1, 3).cuda()
def fwd():
    with torch.no_grad():
        return m(rays, background_rgb=bg).rgb.clone()
def pack(torch_path):
    m._packed_key = None
    orig = fused_step.can_pack_wn
    if torch_path:
        fused_step.can_pack_wn = lambda r: False
    p = m._ensure_packed(dev).clone()
    fused_step.can_pack_wn = orig
    return p
o0 = fwd(); p0 = m._packed.clone()
o1 = fwd()
junk = [torch.randn(1 << 20, device=dev) for _ in range(8)]; del junk
o2 = fwd()
p3 = pack(True); o3 = fwd()
p4 = pack(False); o4 = fwd()
p5 = pack(True); o5 = fwd()
for n, o in (("o1 same pack", o1), ("o2 after junk", o2), ("o3 torch pack", o3), ("o4 wn pack", o4), ("o5 torch pack", o5)):
    print(n, float((o - o0).abs().max()))
for n, p in (("p3", p3), ("p4", p4), ("p5", p5)):
    ne = (p != p0)
    idx = ne.nonzero().flatten()
    print(n, "bytes differing", int(ne.sum()), "first", idx[:4].tolist(), "last", idx[-4:].tolist())
f0, f3 = p0[: p0.numel() // 4 * 4].view(torch.float32), p3[: p3.numel() // 4 * 4].view(torch.float32)
ne = (p0 != p3).view(-1, 4).any(1) if p0.numel() % 4 == 0 else None
idx = ne.nonzero().flatten()
print("float idx differing", idx.numel(), idx[:12].tolist())
for i in idx[:12].tolist():
    print(i, float(f0[i]), float(f3[i]))
blk = 65536
cnt = torch.zeros((f0.numel() + blk - 1) // blk, dtype=torch.long)
for b in range(cnt.numel()):
    cnt[b] = int(ne[b * blk:(b + 1) * blk].sum())
print("per-64K-float block counts", cnt.tolist())
# relative difference where both finite
ok = torch.isfinite(f0) & torch.isfinite(f3) & ne
rel = ((f0 - f3).abs() / f3.abs().clamp_min(1e-9))[ok]
print("rel diff: max", float(rel.max()), "median", float(rel.median()))
