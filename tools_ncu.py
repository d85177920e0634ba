"""Small helpers to summarise ncu outputs (launch lists and .ncu-rep files) into text for profiles/."""
import collections, csv, subprocess, sys


def launches(path, last_n=None):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    if last_n: rows = rows[-last_n:]
    agg = collections.OrderedDict(); tot = 0
    for row in rows:
        n = row['Kernel Name']; v = float(row['Metric Value'].replace(',', ''))
        agg.setdefault(n, [0, 0]); agg[n][0] += v; agg[n][1] += 1; tot += v
    out = [f"total {tot/1e6:.3f} ms over {len(rows)} launches"]
    for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append(f"{v/1e6:10.3f} ms  {c:4d}x  {100*v/tot:5.1f}%  {n[:120]}")
    return "\n".join(out), rows


def details(rep, kid='0'):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'details', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.DictReader(txt.splitlines()))
    want = ['Duration', 'Elapsed Cycles', 'SM Frequency', 'Compute (SM) Throughput', 'Memory Throughput', 'L2 Cache Throughput', 'DRAM Throughput',
            'Registers Per Thread', 'Executed Ipc Active', 'Issue Slots Busy', 'L1/TEX Hit Rate', 'L2 Hit Rate', 'Achieved Occupancy', 'No Eligible',
            'Warp Cycles Per Issued Instruction', 'Dynamic Shared Memory Per Block', 'Grid Size', 'Block Size']
    out = []
    for r in rows:
        if r['ID'] == kid and r['Metric Name'] in want:
            out.append(f"{r['Section Name'][:30]:30s} {r['Metric Name']:40s} {r['Metric Unit']:14s} {r['Metric Value']}")
    return "\n".join(out)


def raw(rep, keys, kid=0):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2 + kid]
    out = []
    for i, h in enumerate(hdr):
        if any(h.endswith(k) for k in keys):
            out.append(f"{h} = {r[i]} {units[i]}")
    return "\n".join(out)


def hot_sass(rep, top=40, kernel_index=0):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    lines = txt.split('\n')
    starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
    a = starts[kernel_index]; b = starts[kernel_index + 1] - 1 if kernel_index + 1 < len(starts) else len(lines)
    rows = list(csv.DictReader(lines[a:b]))
    k = 'Warp Stall Sampling (All Samples)'
    rows2 = [r for r in rows if r.get(k) not in ('', '0', None)]
    tot = sum(float(r[k]) for r in rows2)
    rows2.sort(key=lambda r: -float(r[k]))
    out = [f"{len(rows)} SASS instructions, {tot:.0f} stall samples"]
    for r in rows2[:top]:
        out.append(f"{float(r[k])/tot*100:6.2f}%  exec {r['Instructions Executed']:>10s}  {r['Source'][:110]}")
    return "\n".join(out)


KEYS = ['sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum',
        'lts__t_bytes.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active']

if __name__ == '__main__':
    cmd = sys.argv[1]
    if cmd == 'launches':
        print(launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)[0])
    elif cmd == 'rep':
        kid = int(sys.argv[3]) if len(sys.argv) > 3 else 0
        print(details(sys.argv[2], str(kid))); print(raw(sys.argv[2], KEYS, kid)); print(hot_sass(sys.argv[2], 45, kid))
