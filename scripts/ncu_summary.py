"""Turn ncu artefacts brought back from the GPU box into small text summaries for profiles/.

    python scripts/ncu_summary.py raw   <file.ncu-rep>           # key metrics of every captured launch
    python scripts/ncu_summary.py list  <launches.csv> [anchor]  # per-kernel share of one step
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"== {d.get('Kernel Name', '?')[:100]}")
        for k in KEYS:
            if k in d:
                print(f"   {k:72s} {d[k]:>16s} {units[hdr.index(k)]}")


def launch_list(path, anchor="k_scatter_ones"):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)
            if r.get("Metric Name") == "gpu__time_duration.sum"]
    idx = [i for i, r in enumerate(rows) if anchor in r[0]]
    seg = rows[idx[-3]:idx[-2]] if len(idx) >= 3 else rows
    total = sum(t for _, t in seg)
    agg = collections.OrderedDict()
    for k, t in seg:
        import re
        m = re.search(r"(k_[A-Za-z0-9_]+(?:<[^>]*>)?)", k)
        name = m.group(1)[:60] if m else k.split("(")[0].split("::")[-1][-60:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    print(f"one train step (under ncu: cold cache, serialised): {len(seg)} launches, {total/1e3:.1f} us")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t/1e3:9.1f} us  {100*t/total:5.1f} %  x{c:<3d} {k}")


if __name__ == "__main__":
    {"raw": raw, "list": launch_list}[sys.argv[1]](*sys.argv[2:])
