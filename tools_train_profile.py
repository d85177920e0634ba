import sys, torch
from types import SimpleNamespace
sys.path.insert(0, "/root/repo")
import nrhints_b200 as nb
from nrhints_b200.workload import synthetic_pixel_bundle
dev = torch.device("cuda", 0)
torch.manual_seed(3407)
R = 4096
pb, cam = synthetic_pixel_bundle(R, seed=3407)
pipe = nb.NRHintPipeline(nb.NeuSModelConfig(), nb.RayGeneratorConfig(), nb.CameraModel(**cam), 64).to(dev)
pixels = SimpleNamespace(**{k: v.to(dev) for k, v in vars(pb).items()})
opt = pipe.make_optimizer()
MODE = sys.argv[1] if len(sys.argv) > 1 else "direct"      # direct: NRHintPipeline.train_step; autograd: loss.backward() through the fused node
def step():
    if MODE == "direct":
        pipe.train_step(pixels, global_step=60000, optimizer=opt)
        return
    opt.zero_grad()
    res = pipe(pixels, global_step=60000)
    pipe.get_train_loss_dict(res, pixels)["loss"].backward()
    opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for i in range(5):
    ev[i].record(); step()
ev[5].record(); torch.cuda.synchronize()
print("MODE", MODE, "ms/step", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(5)])
