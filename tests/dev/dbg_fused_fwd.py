import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import nrh_testlib as T
import nrhints_b200 as nb
from oracle import nrh_oracle as orc
cfg = nb.NeuSModelConfig()
m = nb.NeuSHintRenderer(cfg, mlp_impl="auto"); m.load_state_dict(T.make_state("init", cfg)); m.cuda()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rays = orc.synthetic_rays(R, seed=11, crop=500)
dev = {k: v.cuda() for k, v in rays.items()}
out = m(nb.RayBundle(**dev), is_training=True, background_rgb=torch.ones(1, 3).cuda(), global_step=25000)
torch.cuda.synchronize()
print("forward ok", float(out.rgb.mean()))
out.rgb.sum().backward()
torch.cuda.synchronize()
print("backward ok")
