"""Small ragged inference + training-mode forwards (for compute-sanitizer runs; GPU box)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
torch.set_grad_enabled(False)
import nrh_testlib as T
import nrhints_b200 as nb
from nrhints_b200.workload import synthetic_rays
bg = torch.ones(1, 3, device="cuda")
for name, R in (("cfg2_32x128", 301), ("cfg1_64x32", 37), ("outside_16x64", 45), ("maxpoint_16x64", 33)):
    case = T.CASES[name]
    cfg = T.make_config(case)
    m = nb.NeuSHintRenderer(cfg); m.load_state_dict(T.make_state(case["weights"], cfg)); m.cuda()
    rays = nb.RayBundle(**synthetic_rays(R, seed=1, crop=case.get("crop", 300))).to("cuda")
    for training in (False, True):
        out = m(rays, is_training=training, background_rgb=bg, global_step=60000)
        torch.cuda.synchronize()
        print(name, R, training, float(out.rgb.mean()))
