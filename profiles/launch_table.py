"""Print one frame's kernels (time, warp instructions, issue utilisation) from an ncu --metrics CSV."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
by = collections.OrderedDict()
for r in csv.DictReader(lines):
    by.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0][-30:], "grid": r["Grid Size"]})[r["Metric Name"]] = r["Metric Value"]
seq = list(by.values())
idx = [i for i, s in enumerate(seq) if "preprocess" in s["name"]]
tot = 0.0
for s in seq[idx[1]:idx[2]]:
    us = float(s["gpu__time_duration.sum"].replace(",", "")) / 1e3
    tot += us
    extra = ""
    if "smsp__inst_executed.sum" in s:
        extra = f" Minst={float(s['smsp__inst_executed.sum'].replace(',', ''))/1e6:7.1f} issue%={s.get('smsp__issue_active.avg.pct_of_peak_sustained_active', '')}"
    print(f"{s['name']:32s} {s['grid']:16s} us={us:7.1f}{extra}")
print(f"{'sum':32s} {'':16s} us={tot:7.1f}")
