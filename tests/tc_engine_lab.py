"""Developer lab for the tcgen05 SDF engine generations (GPU box): correctness vs the fp64 oracle, kernel time over 524288 points,
and the clock64 timeline of one tile for every generation that covers the requested mode.
usage: python tests/tc_engine_lab.py [gens, e.g. 1,2,3]"""
import sys
import tc_dev
import torch
torch.set_grad_enabled(False)
import nrh_testlib as T
import nrhints_b200 as nb
from oracle import nrh_oracle as orc

gens = [int(g) for g in (sys.argv[1] if len(sys.argv) > 1 else "1,2,3").split(",")]
DBG = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # ablation code (results are wrong with dbg != 0; timings stay valid)
print("dbg", DBG)
cfg = nb.NeuSModelConfig()
sd = T.make_state("sharp", cfg)
m = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m.load_state_dict(sd); m.cuda()
g = torch.Generator().manual_seed(1)
pts = torch.cat([(torch.rand(1000, 3, generator=g) - 0.5) * 2.6, 4.5 * torch.nn.functional.normalize(torch.randn(77, 3, generator=g), dim=-1)])
ocfg = orc.OracleConfig.from_model_config(cfg)
w64 = orc.sdf_mlp(orc.effective_weights(sd, torch.float64), pts.double(), ocfg, want_feat=True, want_grad=True)
big = (torch.rand(4096 * 128, 3, device="cuda") - 0.5) * 2
modes = [(False, False), (True, False), (True, True)]
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
for gen in gens:
    for wg, wf in modes:
        tc_dev.configure(gen=gen, dbg=DBG)
        sdf, grad, feat = m.sdf_query(pts.cuda(), want_grad=wg, want_feat=wf)
        torch.cuda.synchronize()
        e = ["sdf %.2e" % float((sdf.cpu().double() - w64["sdf"][:, 0]).abs().max())]
        if grad is not None: e.append("grad %.2e" % float((grad.cpu().double() - w64["grad"]).abs().max()))
        if feat is not None: e.append("feat %.2e" % float((feat.cpu().double() - w64["feat"]).abs().max()))
        for _ in range(2): m.sdf_query(big, want_grad=wg, want_feat=wf)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): m.sdf_query(big, want_grad=wg, want_feat=wf)
        b.record(); torch.cuda.synchronize()
        print(f"gen {gen} grad={wg} feat={wf}: {a.elapsed_time(b) / 5:.3f} ms / 524288 points; err vs fp64 {e}", flush=True)
    # timeline of the sdf-only pass (generation 2/3 layout: 16 slots per gemm for the MMA thread, 16 per layer for the epilogue)
    if gen in (2, 3):
        tc_dev.configure(gen=gen, dbg=DBG, tlog=buf)
        m.sdf_query(big); torch.cuda.synchronize()
        buf.zero_(); m.sdf_query(big); torch.cuda.synchronize()
        t = buf.cpu().numpy(); base = t[t > 0].min()
        print(f"gen {gen} MMA thread per gemm: K-half 0 (N=256) | K-half 1 N-half 0 | K-half 1 N-half 1: [group start, first operand ready, group issued]")
        for gi in range(8):
            r = t[gi * 16: gi * 16 + 12] - base
            print(f" g{gi}: " + " | ".join(f"{r[i*3]:6d} {r[i*3+1]:6d} {r[i*3+2]:6d}" for i in ((0, 2, 3) if gi else (0,))))
        starts = [int(t[gi * 16] - base) for gi in range(8)]
        print(f" layer period (gemm start to gemm start, g2..g7): {[starts[i+1]-starts[i] for i in range(2, 7)]}")
        if gen == 2:
            print(" epilogue warp per layer [wait h0 start, h0 ready, step0 done | step1 done | wait h1 start, h1 ready, step2 done | step3 done]")
            for l in range(7):
                r = t[256 + l * 16: 256 + l * 16 + 12] - base
                print(f" l{l}: {r[0]:6d} {r[1]:6d} {r[2]:6d} | {r[5]:6d} | {r[6]:6d} {r[7]:6d} {r[8]:6d} | {r[11]:6d}")
        else:
            print(" epilogue per layer, team 0 then team 1: [wait start, half ready, step0..3 done]")
            for l in range(7):
                r = t[256 + l * 16: 256 + l * 16 + 16] - base
                print(f" l{l}: T0 " + " ".join(f"{r[i]:6d}" for i in range(6)) + "   T1 " + " ".join(f"{r[8+i]:6d}" for i in range(6)))
        tc_dev.configure(gen=gen)
tc_dev.configure(gen=1)
