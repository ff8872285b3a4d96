"""Summarise ncu outputs: `launches` CSV (per-kernel device time) and `raw` page of a full capture."""
import collections, csv, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        agg.setdefault(row["Kernel Name"].split("(")[0][:48], []).append(float(row["Metric Value"].replace(",", "")) / 1e3)
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print(f"{k:50s} n={len(v):3d} mean={sum(v)/len(v):9.1f} us  share={100*sum(v)/tot:5.1f}%")

TERSE = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum"]
STALLS = "smsp__average_warps_issue_stalled_"

def terse(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        st = sorted(((float(v), k[len(STALLS):].split("_per_")[0]) for k, v in d.items() if k.startswith(STALLS) and k.endswith("ratio") and "selected" not in k), reverse=True)[:4]
        print(d["Kernel Name"].split("(")[0][:32], "|", " ".join(f"{k.split('.')[0].split('__')[-1]}={d[k]}" for k in TERSE if k in d),
              "| stalls:", " ".join(f"{n}={v:.1f}" for v, n in st))

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_lsu",
        "smsp__inst_executed_pipe_fma", "smsp__inst_executed_pipe_fmaheavy", "smsp__inst_executed_pipe_fmalite",
        "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_alu_cycles_active.avg.pct", "sm__pipe_xu_cycles_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "smsp__average_warps_issue_stalled"]

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index("Kernel Name")][:70])
        for h, u, v in zip(hdr, units, r):
            if any(h.startswith(k) for k in KEYS):
                print(f"  {h:86s} {v:>16s} {u}")

if __name__ == "__main__":
    {"launches": launches, "raw": raw, "terse": terse}[sys.argv[1]](sys.argv[2])
