#!/usr/bin/env python
"""Per-kernel census of the Blackwell-specific SASS opcodes in the in-tree library (cuobjdump -sass): UTC*MMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UBLKCP (cp.async.bulk, non-tensor TMA), UTMALDG / UTMASTG (cp.async.bulk.tensor), UTCBAR (tcgen05.commit), SYNCS
(mbarrier), HMMA (legacy mma.sync -- must be absent).  Runs without a GPU.   python tools_opcode_census.py > profiles/<round>_opcode_census.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent / "nrhints_b200" / "csrc" / "libnrhints_b200.so"
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "MUFU", "FFMA"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = cur.replace("(anonymous namespace)::", "").replace("void ", "").replace("nrh::", "")
            cur = re.sub(r"\(.*", "", cur)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
            counts[cur]["total"] += 1
    print(f"SASS opcode census of {LIB.name} (sm_100a), {len(counts)} kernels; columns: " + " ".join(OPS) + " | total instructions")
    for k, c in sorted(counts.items(), key=lambda kv: -(kv[1]["UTCHMMA"] + kv[1]["UTCQMMA"])):
        print(f"{k[:70]:70s} " + " ".join(f"{c[o]:6d}" for o in OPS) + f" | {c['total']:7d}")
    tc = [k for k, c in counts.items() if c["UTCHMMA"] + c["UTCQMMA"] + c["UTCIMMA"] > 0]
    print(f"\nkernels issuing tcgen05.mma: {len(tc)}: " + ", ".join(tc))
    print("legacy HMMA (mma.sync) instructions anywhere:", sum(c["HMMA"] for c in counts.values()))
    print("tensor-map TMA (UTMALDG/UTMASTG) instructions anywhere:", sum(c["UTMALDG"] + c["UTMASTG"] for c in counts.values()),
          "(weights / operand images are pre-swizzled and moved by plain bulk copies, UBLKCP)")


if __name__ == "__main__":
    main()
