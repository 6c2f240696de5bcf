"""Hottest SASS regions of one kernel in an ncu report: executed warp instructions per run of consecutive addresses,
with the opcode mix of each.   python tools/ncu_sass_hot.py <report.ncu-rep> <kernel regex> [occurrence] [n]"""
import collections, csv, io, re, subprocess, sys

rep, want = sys.argv[1], sys.argv[2]
occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:k_"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
sections, cur, hdr = [], None, None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
        hdr = None
    elif r and r[0] == "Address":
        hdr = r
    elif hdr and cur is not None and len(r) == len(hdr):
        cur["rows"].append(dict(zip(hdr, r)))
sections = [s for s in sections if re.search(want, s["name"]) and s["rows"]]
sec = sections[occ]
v = sec["rows"]
tot = sum(int(d["Instructions Executed"] or 0) for d in v)
smp = sum(int(d["# Samples"] or 0) for d in v)
print(f"# {sec['name'][:70]}: {len(v)} SASS instructions, {tot} executed (warp level), {smp} samples")
# group in windows of 64 static instructions
W = 64
for i in range(0, len(v), W):
    blk = v[i:i + W]
    c = sum(int(d["Instructions Executed"] or 0) for d in blk)
    s = sum(int(d["# Samples"] or 0) for d in blk)
    if c < tot * 0.015 and s < smp * 0.015:
        continue
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", d["Source"]).split()[0].split(".")[0] for d in blk)
    thr = sum(int(d["Thread Instructions Executed"] or 0) for d in blk) / max(c, 1)
    print(f"{i:6d} {100 * c / tot:5.1f}% inst {100 * s / max(smp,1):5.1f}% samples thr {thr:4.1f}  " + " ".join(f"{k}:{n}" for k, n in ops.most_common(7)))
if len(sys.argv) > 6:
    a, b = int(sys.argv[5]), int(sys.argv[6])
    for i in range(a, b):
        d = v[i]
        print(f"{i:6d} {int(d['Instructions Executed'] or 0):10d} {float(d['Avg. Threads Executed'] or 0):5.1f} {int(d['# Samples'] or 0):6d}  {d['Source'][:90]}")
