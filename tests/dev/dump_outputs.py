"""Developer tool (GPU box): render a fixed set of cases and save every output tensor, to compare two builds of the library bitwise.
usage: dump_outputs.py out.pt   |   dump_outputs.py --compare a.pt b.pt"""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
if sys.argv[1] == "--compare":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    bad = 0
    diffs = []
    for k in a:
        if k.startswith("time"):
            continue
        if not torch.equal(a[k], b[k]):
            bad += 1
            diffs.append((float((a[k].float() - b[k].float()).abs().max()), k))
    for dmax, k in sorted(diffs, reverse=True)[:12]:
        print("DIFF", k, dmax)
    print(f"{len(a)} tensors compared, {bad} differ;", {k: (a[k], b[k]) for k in a if k.startswith("time")})
    sys.exit(1 if bad else 0)
import numpy as np
import nrh_testlib as T
import nrhints_b200 as nb
from nrhints_b200.workload import synthetic_rays
torch.set_grad_enabled(False)
res = {}
bg = torch.ones(1, 3, device="cuda")


def run(tag, cfg, weights, R, training=False, seed=0, crop=800):
    m = nb.NeuSHintRenderer(cfg, mlp_impl="auto"); m.load_state_dict(T.make_state(weights, cfg)); m.cuda()
    rays = nb.RayBundle(**synthetic_rays(R, seed=seed, crop=crop)).to("cuda")
    torch.manual_seed(1234)
    out = m(rays, is_training=training, background_rgb=bg, global_step=60000, return_extras=True)
    torch.cuda.synchronize()
    for k, v in out.as_dict().items():
        if v is not None:
            res[f"{tag}.{k}"] = v.detach().cpu()
    return m, rays


m, rays = run("sharp4096", nb.NeuSModelConfig(), "sharp", 4096, seed=4242)
run("init1000", nb.NeuSModelConfig(), "init", 1000, seed=7, crop=300)
run("train777", nb.NeuSModelConfig(), "sharp", 777, training=True, seed=9, crop=300)
for name in ("cfg1_64x32", "shadowonly_16x64", "nohint_16x64", "maxpoint_16x64", "outside_16x64"):
    case = T.CASES[name]
    run(name, T.make_config(case), case["weights"], 37 if name != "cfg1_64x32" else 64, seed=case["ray_seed"], crop=case.get("crop", 300))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    m(rays, background_rgb=bg)
ts = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); m(rays, background_rgb=bg); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
res["time_forward_ms"] = torch.tensor(float(np.mean(ts)))
torch.save(res, sys.argv[1])
print("saved", len(res), "tensors; forward", float(np.mean(ts)), "ms")
