"""Test micro-F1 over the completed repeats of GLASSTest.py logs (`end: epoch ..., val V, tst T` lines): mean, standard
error, count -- for the reference's own CPU logs and this repo's GPU logs alike.
    python scripts/score_summary.py LOG [LOG ...]"""
import math, re, sys
for path in sys.argv[1:]:
    tst, val, secs = [], [], []
    for line in open(path, errors="replace"):
        m = re.search(r"end: epoch (\d+), train time ([\d.]+) s, val ([\d.]+), tst ([\d.]+)", line)
        if m:
            secs.append(float(m.group(2))); val.append(float(m.group(3))); tst.append(float(m.group(4)))
    n = len(tst)
    if not n:
        print(f"{path}: no completed repeats")
        continue
    mean = sum(tst) / n
    se = math.sqrt(sum((t - mean) ** 2 for t in tst) / max(n - 1, 1) / n)
    print(f"{path}: repeats {n}  tst {mean:.4f} +- {se:.4f} (SE)  val {sum(val) / n:.4f}  train time {sum(secs) / n:.2f} s/run")
