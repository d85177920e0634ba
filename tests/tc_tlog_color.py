"""Developer timeline of the reflectance tcgen05 kernel (GPU box)."""
import tc_dev
import torch
torch.set_grad_enabled(False)
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
tc_dev.configure(tlog=buf)
import nrh_testlib as T
import nrhints_b200 as nb
from oracle import nrh_oracle as orc
cfg = nb.NeuSModelConfig(); sd = T.make_state("init", cfg)
m = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m.load_state_dict(sd); m.cuda()
b = nb.RayBundle(**orc.synthetic_rays(4096, seed=3407)).to("cuda")
m(b, background_rgb=torch.ones(1, 3).cuda()); torch.cuda.synchronize()
buf.zero_(); m(b, background_rgb=torch.ones(1, 3).cuda()); torch.cuda.synchronize()
t = buf.cpu().numpy()[256:]
base = t[t > 0].min()
print("color kernel, block 0, third tile (cycles):")
print(" epilogue warp: stage_start %d stage_done %d" % (t[64] - base, t[65] - base))
for gi in range(4):
    print(f" gemm {gi}: MMA wait_a0 {t[gi*8]-base:6d} a0_ready {t[gi*8+1]-base:6d} commit {t[gi*8+2]-base:6d} | epilogue acc_ready {t[66+gi*2]-base:6d} done {t[67+gi*2]-base:6d}")
