"""Three inference forwards of 4096 x 128 (for ncu captures of the non-MLP kernels; GPU box)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
torch.set_grad_enabled(False)
import nrhints_b200 as nb
from nrhints_b200.workload import synthetic_rays
torch.manual_seed(3407)
m = nb.NeuSHintRenderer(nb.NeuSModelConfig()).cuda()
rays = nb.RayBundle(**synthetic_rays(4096, seed=3407)).to("cuda")
bg = torch.ones(1, 3, device="cuda")
for _ in range(3):
    m(rays, background_rgb=bg)
torch.cuda.synchronize()
