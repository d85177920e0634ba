import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import nrh_testlib as T
import nrhints_b200 as nb
from nrhints_b200 import fused_step
dev = torch.device("cuda", 0)
cfg = nb.NeuSModelConfig()
m = nb.NeuSHintRenderer(cfg, mlp_impl="auto"); m.load_state_dict(T.make_state("sharp", cfg)); m.cuda()
def pack(torch_path):
    m._packed_key = None
    orig = fused_step.can_pack_wn
    if torch_path:
        fused_step.can_pack_wn = lambda r: False
    p = m._ensure_packed(dev).clone()
    fused_step.can_pack_wn = orig
    torch.cuda.synchronize()
    return p, (m._wn_scratch.clone() if getattr(m, "_wn_scratch", None) is not None else None)
seq = [False, False, True, False, True, True, False]
ps = [pack(t) for t in seq]
for i in range(1, len(ps)):
    print("pack", i, "torch" if seq[i] else "wn", "bytes differing vs pack 0:", int((ps[i][0] != ps[0][0]).sum()),
          "vs previous:", int((ps[i][0] != ps[i - 1][0]).sum()),
          "scratch differing vs pack 0:", int((ps[i][1] != ps[0][1]).sum()) if ps[i][1] is not None and ps[0][1] is not None else None)
