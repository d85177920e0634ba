"""Developer ablations of the tcgen05 SDF kernel (GPU box): usage tc_ablate.py [dbg code]"""
import sys
import tc_dev
import torch
import nrh_testlib as T
import nrhints_b200 as nb
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
tc_dev.configure(gen=1, dbg=mode)
cfg = nb.NeuSModelConfig(); sd = T.make_state("init", cfg)
m = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m.load_state_dict(sd); m.cuda()
pts = (torch.rand(4096 * 128, 3, device="cuda") - 0.5) * 2
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
print(f"dbg={mode}  sdf-only {t(lambda: m.sdf_query(pts)):.3f} ms   grad {t(lambda: m.sdf_query(pts, want_grad=True)):.3f} ms   grad+feat {t(lambda: m.sdf_query(pts, want_grad=True, want_feat=True)):.3f} ms   (524288 pts)")
