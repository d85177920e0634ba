import os, sys
import torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
os.environ["NRH_TC_TLOG"] = hex(buf.data_ptr())
import nrh_testlib as T
import nrhints_b200 as nb
cfg = nb.NeuSModelConfig(); sd = T.make_state("init", cfg)
m = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m.load_state_dict(sd); m.cuda()
pts = (torch.rand(4096 * 128, 3, device="cuda") - 0.5) * 2
grad = len(sys.argv) > 1 and sys.argv[1] == "grad"
if len(sys.argv) > 2: os.environ["NRH_TC_DEBUG"] = sys.argv[2]
m.sdf_query(pts, want_grad=grad); torch.cuda.synchronize()
buf.zero_(); m.sdf_query(pts, want_grad=grad); torch.cuda.synchronize()
t = buf.cpu().numpy()
base = t[t > 0].min()
print("mode", "grad" if grad else "sdf-only", "dbg", os.environ.get("NRH_TC_DEBUG", "0"))
print("MMA thread, per gemm gi: [wait_a_start, a_ready, chunk_issued] x4 chunks, then acc commit  (cycles from first stamp)")
for gi in range(8):
    r = t[gi * 16: gi * 16 + 13] - base
    print(f" g{gi}: " + " | ".join(f"{r[c*3]:6d} {r[c*3+1]:6d} {r[c*3+2]:6d}" for c in range(4)) + f" | commit {r[12]:6d}")
print("epilogue warp 2, per layer l: wait_acc_start, acc_ready, then per chunk [ld_done, math_done, published]")
for l in range(7):
    r = t[128 + l * 16: 128 + l * 16 + 14] - base
    print(f" l{l}: {r[0]:6d} {r[1]:6d} | " + " | ".join(f"{r[2+c*3]:6d} {r[3+c*3]:6d} {r[4+c*3]:6d}" for c in range(4)))
