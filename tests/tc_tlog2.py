"""Developer timeline of the second-generation tcgen05 SDF kernel (clock64 stamps of block 0, third tile); GPU box."""
import os, sys
import torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
os.environ["NRH_TC_TLOG"] = hex(buf.data_ptr()); os.environ["NRH_TC_GEN"] = "2"
if len(sys.argv) > 1: os.environ["NRH_TC_DEBUG"] = sys.argv[1]
print("dbg", os.environ.get("NRH_TC_DEBUG", "0"))
import nrh_testlib as T
import nrhints_b200 as nb
cfg = nb.NeuSModelConfig(); sd = T.make_state("init", cfg)
m = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m.load_state_dict(sd); m.cuda()
pts = (torch.rand(4096 * 128, 3, device="cuda") - 0.5) * 2
m.sdf_query(pts); torch.cuda.synchronize()
buf.zero_(); m.sdf_query(pts); torch.cuda.synchronize()
t = buf.cpu().numpy(); base = t[t > 0].min()
print("MMA thread per gemm: groups K-half 0 (N=256), K-half 1 / N-half 0, K-half 1 / N-half 1: [group start, first operand ready, group issued]")
for gi in range(8):
    r = t[gi * 16: gi * 16 + 12] - base
    print(f" g{gi}: " + " | ".join(f"{r[i*3]:6d} {r[i*3+1]:6d} {r[i*3+2]:6d}" for i in ((0, 2, 3) if gi else (0,))))
print("epilogue warp: per layer [wait h0 start, h0 ready, step0 done | step1 done | wait h1 start, h1 ready, step2 done | step3 done]")
for l in range(7):
    r = t[256 + l * 16: 256 + l * 16 + 12] - base
    print(f" l{l}: {r[0]:6d} {r[1]:6d} {r[2]:6d} | {r[5]:6d} | {r[6]:6d} {r[7]:6d} {r[8]:6d} | {r[11]:6d}")
