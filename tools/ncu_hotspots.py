"""Source-view summary of one `ncu --set full --import-source on` report (exported with
`ncu -i X.ncu-rep --page source --csv > X.csv`): stall samples by loop (SASS lines grouped by execution count), the
stall reasons of the hottest loops, and the hottest single instructions.  Writes markdown to stdout.

    python tools/ncu_hotspots.py X.csv "title" > profiles/r02_..._hotspots.md
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
kernel = rows[0][1] if len(rows[0]) > 1 else ""
hdr = rows[1]
ix_s, ix_n, ix_i = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((int(r[ix_n]), int(r[ix_i]), k, r[ix_s].strip(), r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
toti = sum(d[1] for d in data)
print(f"# {title}\n\nkernel `{kernel}`; {tot} stall samples, {toti} warp instructions, {len(data)} SASS lines.\n")
print("## Samples by loop (SASS lines grouped by how often they executed)\n")
print("| executions per line | SASS lines | share of samples | share of instructions | top stall reasons |\n|---:|---:|---:|---:|---|")
groups = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
for n, ie, k, s, r in data:
    g = groups[ie]
    g[0] += n; g[1] += ie; g[2] += 1
    for i, h in stall_cols:
        if r[i].isdigit():
            g[3][h] += int(r[i])
for ie, (n, i, c, st) in sorted(groups.items(), key=lambda x: -x[1][0])[:8]:
    reasons = ", ".join(f"{h[6:]} {100 * v / max(n, 1):.0f} %" for h, v in st.most_common(4))
    print(f"| {ie} | {c} | {100 * n / tot:.1f} % | {100 * i / toti:.1f} % | {reasons} |")
print("\n## Opcode mix of the hottest loop\n")
hot = max(groups.items(), key=lambda x: x[1][0])[0]
ops = collections.Counter(); ops_s = collections.Counter()
for n, ie, k, s, r in data:
    if ie == hot:
        op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
        ops[op] += 1; ops_s[op] += n
tn, ts = sum(ops.values()), sum(ops_s.values())
print("| opcode | instructions per pass | share | share of the loop's samples |\n|---|---:|---:|---:|")
for op, v in ops.most_common(12):
    print(f"| {op} | {v} | {100 * v / tn:.1f} % | {100 * ops_s[op] / max(ts, 1):.1f} % |")
print("\n## Hottest instructions\n")
print("| share of samples | executions | SASS | main stall reason |\n|---:|---:|---|---|")
for n, ie, k, s, r in sorted(data, key=lambda d: -d[0])[:15]:
    st = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for i, h in stall_cols), reverse=True)[0]
    print(f"| {100 * n / tot:.2f} % | {ie} | `{s[:70]}` | {st[1][6:]} |")
