"""Developer timeline of the tcgen05 SDF kernel (clock64 stamps of block 0, third tile); run on the GPU box.
usage: tc_tlog.py [sdf|grad] [NRH_TC_DEBUG value]"""
import os, sys
import tc_dev
import torch
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
import nrh_testlib as T
import nrhints_b200 as nb
cfg = nb.NeuSModelConfig(); sd = T.make_state("init", cfg)
m = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m.load_state_dict(sd); m.cuda()
pts = (torch.rand(4096 * 128, 3, device="cuda") - 0.5) * 2
grad = len(sys.argv) > 1 and sys.argv[1] in ("grad", "fine")
feat = len(sys.argv) > 1 and sys.argv[1] == "fine"
dbg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
fmask = int(os.environ.get("NRH_TC_FMASK", "0xFF"), 0)
tc_dev.configure(gen=1, dbg=dbg, fmask=fmask, tlog=buf)
m.sdf_query(pts, want_grad=grad, want_feat=feat); torch.cuda.synchronize()
buf.zero_(); m.sdf_query(pts, want_grad=grad, want_feat=feat); torch.cuda.synchronize()
t = buf.cpu().numpy()
base = t[t > 0].min()
print("mode", "grad" if grad else "sdf-only", "dbg", str(dbg))
print("MMA thread, per gemm gi: [a_ready, issued] x8 sub-chunks (relative to the wait start of sub-chunk 0), then acc commit")
for gi in range(8):
    r = t[gi * 32: gi * 32 + 25] - base
    n = 2 if gi == 0 else 8
    print(f" g{gi}: start {r[0]:6d} | " + " | ".join(f"{r[c*3+1]-r[0]:5d} {r[c*3+2]-r[0]:5d}" for c in range(n)) + f" | commit {r[24]-r[0]:5d}")
print("epilogue warp 2, per layer l: wait_acc_start, acc_ready, then per sub-chunk [ld_done, math_done, published] relative to acc_ready")
for l in range(7):
    r = t[256 + l * 32: 256 + l * 32 + 26] - base
    print(f" l{l}: {r[0]:6d} {r[1]:6d} | " + " | ".join(f"{r[2+c*3]-r[1]:5d} {r[3+c*3]-r[1]:5d} {r[4+c*3]-r[1]:5d}" for c in range(8)))
cm = [int(t[gi * 32 + 24] - base) for gi in range(8)]
wa = [int(t[gi * 32 + 1] - t[gi * 32 + 0]) for gi in range(2, 7)]
ew = [int(t[256 + l * 32 + 1] - t[256 + l * 32]) for l in range(2, 7)]
print(f"SUMMARY cluster={os.environ.get('NRH_TC_CLUSTER','1')} dbg={dbg} period/gemm {(cm[6]-cm[2])/4:.0f}  "
      f"MMA first-sub-chunk wait {sum(wa)/len(wa):.0f}  epilogue accumulator wait {sum(ew)/len(ew):.0f}")

# every epilogue warp (lane 0): publish stamps of forward layer 2 and reverse layer 5, relative to the earliest accumulator wake-up
for name, off in (("forward layer 2", 512), ("reverse layer 5", 768)):
    blk = t[off: off + 256].reshape(16, 16)
    if not (blk[:, 8] > 0).any():
        continue
    t0 = blk[:, 8][blk[:, 8] > 0].min()
    print(f"{name}: per epilogue warp (scheduler = warp % 4, column group = warp // 4): accumulator wake-up, then publish of sub-chunks 0..7, relative to the first wake-up")
    for w in range(16):
        print(f"  warp {w:2d} (sched {w % 4}, gq {w // 4}): {int(blk[w, 8] - t0):5d} | " + " ".join(f"{int(blk[w, c] - t0):5d}" for c in range(8)))
    last = blk[:, :8].max(axis=0) - t0
    first = blk[:, :8].min(axis=0) - t0
    print("  first / last warp per sub-chunk:", " ".join(f"{int(a)}/{int(b)}" for a, b in zip(first, last)))
