import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import nrh_testlib as T
import nrhints_b200 as nb
from nrhints_b200 import fused_step
dev = torch.device("cuda", 0)
cfg = nb.NeuSModelConfig()
for kind in ("init", "sharp"):
    m = nb.NeuSHintRenderer(cfg, mlp_impl="auto"); m.load_state_dict(T.make_state(kind, cfg)); m.cuda()
    m._ensure_packed(dev)
    torch.cuda.synchronize()
    w = m._wn_scratch.view(torch.float32)
    off = 0
    for i, lin in enumerate(fused_step.wn_layers(m)):
        e = lin.effective_weight().detach().reshape(-1)
        mine = w[off:off + e.numel()]
        nd = int((mine != e).sum())
        if nd:
            idx = (mine != e).nonzero().flatten()
            rows = sorted(set((idx // lin.weight_v.shape[1]).tolist()))
            print(kind, "layer", i, tuple(lin.weight_v.shape), "differing", nd, "rows", rows[:10], "n rows", len(rows),
                  "g of row", float(lin.weight_g.reshape(-1)[rows[0]]), "norm", float(lin.weight_v[rows[0]].norm()))
        off += e.numel()
    print(kind, "done")
