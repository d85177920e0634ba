"""Sustained SM clock / power under the dominant kernel: loops the fine-pass SDF kernel for a few seconds while nvidia-smi samples."""
import subprocess, sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import nrhints_b200 as nb
m = nb.NeuSHintRenderer(nb.NeuSModelConfig(), mlp_impl="auto").cuda()
pts = (torch.rand(524288, 3, device="cuda") - 0.5) * 1.6
for mode, kw in (("fine fwd+feat+grad", dict(want_grad=True, want_feat=True)), ("sdf only", dict())):
    for _ in range(3):
        m.sdf_query(pts, **kw)
    torch.cuda.synchronize()
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50"],
                         stdout=subprocess.PIPE, text=True)
    t0 = time.time()
    n = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            m.sdf_query(pts, **kw)
        n += 20
        torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    p.terminate()
    out = p.stdout.read().strip().splitlines()
    rows = [l.split(",") for l in out if l.count(",") == 2]
    clk = sorted(float(r[0]) for r in rows[len(rows) // 4:])
    pw = sorted(float(r[1]) for r in rows[len(rows) // 4:])
    cap = sum(1 for r in rows if "Active" in r[2] and "Not" not in r[2])
    print(mode, "ms/launch", round(a.elapsed_time(b) / n, 4), "samples", len(rows), "clk median", clk[len(clk) // 2], "min", clk[0], "max", clk[-1],
          "power median", pw[len(pw) // 2], "max", pw[-1], "power-cap samples", cap)
    # first launch after idle (burst clocks)
    time.sleep(2.0)
    a.record(); m.sdf_query(pts, **kw); b.record(); torch.cuda.synchronize()
    print(mode, "single launch after 2 s idle: ms", round(a.elapsed_time(b), 4))
