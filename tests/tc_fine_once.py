"""One fine-pass launch (sdf + gradient + features over 524288 points) for ncu captures (GPU box).  usage: tc_fine_once.py [dev dbg]"""
import sys
if len(sys.argv) > 1:
    import tc_dev
    tc_dev.configure(gen=1, dbg=int(sys.argv[1]))
else:
    sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
torch.set_grad_enabled(False)
import nrh_testlib as T
import nrhints_b200 as nb
cfg = nb.NeuSModelConfig(); sd = T.make_state("init", cfg)
m = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m.load_state_dict(sd); m.cuda()
pts = (torch.rand(4096 * 128, 3, device="cuda") - 0.5) * 2
for _ in range(3):
    m.sdf_query(pts, want_grad=True, want_feat=True)
torch.cuda.synchronize()
