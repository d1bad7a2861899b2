"""Rewrite the "Trained scores" table of profiles/README.md from the 30-seed logs in profiles/r02/."""
import math, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stats(path):
    tst, secs = [], []
    for line in open(path, errors="replace"):
        m = re.search(r"end: epoch (\d+), train time ([\d.]+) s, val ([\d.]+), tst ([\d.]+)", line)
        if m:
            tst.append(float(m.group(4)))
            secs.append(float(m.group(2)))
    n = len(tst)
    mean = sum(tst) / n
    return n, mean, math.sqrt(sum((t - mean) ** 2 for t in tst) / max(n - 1, 1) / n), sum(secs) / n


paper = {"density": 0.930, "cut_ratio": 0.935, "component": 1.000, "coreness": 0.840}
rows = ["| dataset | this repo, 1 x B200 (seeds) | unmodified reference, CPU (seeds) | difference (pt) | SE of the difference (pt) | paper | train time per run (B200 / CPU) |",
        "|---|---|---|---|---|---|---|"]
worst = 0.0
for ds in ("density", "cut_ratio", "component", "coreness"):
    a = stats(os.path.join(ROOT, "profiles", "r02", f"glasstest_repeat30_{ds}.log"))
    b = stats(os.path.join(ROOT, "profiles", "r02", f"reference_cpu_repeat30_{ds}.log"))
    se = math.sqrt(a[2] ** 2 + b[2] ** 2)
    worst = max(worst, abs(a[1] - b[1]) / se)
    rows.append(f"| {ds} | {a[1]:.4f} +- {a[2]:.4f} ({a[0]}) | {b[1]:.4f} +- {b[2]:.4f} ({b[0]}) | {100 * (a[1] - b[1]):+.2f} | "
                f"{100 * se:.2f} | {paper[ds]:.3f} | {a[3]:.2f} s / {b[3]:.0f} s |")
text = "\n".join(rows) + f"""

Seeds are the reference's own (`(1 << repeat) - 1`, split seeded with 0).  The largest difference is {worst:.1f} standard
errors of the difference (coreness, in this repo's favour and towards the paper's 0.840; its validation scores go the
other way, 0.875 against 0.880, so this is read as seed noise on one fixed 55-subgraph test split, not as a defect of the
reference or a gain of this repo); 30 seeds resolve about 0.8 pt on these 55-63-subgraph test splits (one flipped prediction =
1.6-1.8 pt), so the 0.5 pt bar of the north star can only be met statistically, not seed by seed -- which `--use_one` makes
impossible for any two implementations (DESIGN.md section 2, "Degenerate `--use_one` inputs").  The CPU columns are the unmodified
reference run in the build container (46-202 s per run on 2 threads, `PYTHONPATH=oracle/pyg_shim`).  Logs: `r02/glasstest_repeat30_*.log`,
`r02/reference_cpu_repeat30_*.log`; `python scripts/score_table.py` rewrites this table from them."""
p = os.path.join(ROOT, "profiles", "README.md")
s = open(p).read()
s = re.sub(r"<!-- scores -->.*?<!-- /scores -->", lambda m: "<!-- scores -->\n" + text + "\n<!-- /scores -->", s, flags=re.S)
open(p, "w").write(s)
print(text)
