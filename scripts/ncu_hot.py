"""Hot SASS instructions of one kernel from an ncu report (stall samples per instruction, with stall reasons).

    python scripts/ncu_hot.py <file.ncu-rep> <kernel regex> [launch index] [top N]
"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                      f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
print(rows[h - 1][:2])
ia, isrc, isa = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
reasons = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
data = []
for r in rows[h + 1:]:
    if len(r) <= isa:
        continue
    try:
        s = int(r[isa] or 0)
    except ValueError:
        continue
    rs = {c: int((r[hdr.index(c)] if hdr.index(c) < len(r) else 0) or 0) for c in reasons}
    data.append((r[isrc], s, rs))
tot = sum(d[1] for d in data)
print("total samples", tot, "instructions", len(data))
agg = {c: sum(d[2][c] for d in data) for c in reasons}
print("by reason:", {k: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(range(len(data)), key=lambda i: -data[i][1])[:top_n]
for i in sorted(top):
    src, s, rs = data[i]
    main = max(rs.items(), key=lambda kv: kv[1])
    print(f"{i:5d} {s:6d} {100 * s / tot:5.1f}%  {main[0][6:]:12s} {src.strip()[:100]}")
