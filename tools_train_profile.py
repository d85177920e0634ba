import sys, torch
sys.path.insert(0, "/root/repo")
import nrhints_b200 as nb
from nrhints_b200.workload import synthetic_rays
dev = torch.device("cuda", 0)
torch.manual_seed(3407)
m = nb.NeuSHintRenderer(nb.NeuSModelConfig()).to(dev)
opt = nb.FlatAdam(m.parameters(), lr=5e-4)
R = 4096
rays = nb.RayBundle(**synthetic_rays(R, seed=3407)).to(dev)
bg = torch.ones(1, 3, device=dev); gt = torch.rand(R, 3, device=dev)
def step():
    opt.zero_grad()
    out = m(rays, is_training=True, background_rgb=bg, global_step=60000)
    nb.train_loss_dict(out, gt, 0.1)["loss"].backward()
    opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
