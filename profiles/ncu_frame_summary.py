"""Summarise an `ncu --set full` capture of profiles/prof_frame.py into the two committed files:

    python profiles/ncu_frame_summary.py gpurun_out/r02b_frame.ncu-rep profiles/r02_frame_ncu.txt profiles/r02_kernel_dram.json

The text file lists every captured kernel (time, warp instructions, issue %, occupancy limits, DRAM bytes, pipe
utilisation, top stalls); the JSON sums the launches of a kernel family and is what bench.py quotes as (static) traffic.
"""
import collections, csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def main(rep, txt, js, note):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    fam = collections.OrderedDict()
    with open(txt, "w") as f:
        f.write(f"ncu --set full --clock-control none, one C3 frame of {note} (profiles/prof_frame.py), stage by stage; cold-cache, serialised.\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = d["Kernel Name"].split("(")[0]
            f.write(f"== {name}   grid {d['Grid Size']} block {d['Block Size']}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"   {k:82s} {d[k]:>18s} {u[k]}\n")
            st = sorted(((float(v.replace(',', '')), k[len(STALL):].split("_per_")[0]) for k, v in d.items()
                         if k.startswith(STALL) and k.endswith("warp_active.ratio") is False and k.endswith(".ratio") and "not_issued" not in k and "selected" not in k), reverse=True)[:6]
            f.write("   stalls per issue: " + ", ".join(f"{n}={v:.2f}" for v, n in st) + "\n")
            num = lambda k: float(d[k].replace(",", "")) * SCALE.get(u[k], 1.0)
            base = name.replace("void ", "").split("<")[0]
            e = fam.setdefault(base, {"kernel": name, "launches": 0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "time_us": 0.0, "warp_inst": 0.0})
            e["launches"] += 1
            e["dram_bytes_read"] += num("dram__bytes_read.sum")
            e["dram_bytes_write"] += num("dram__bytes_write.sum")
            e["time_us"] += num("gpu__time_duration.sum")
            e["warp_inst"] += num("smsp__inst_executed.sum")
    for e in fam.values():
        e["dram_bytes"] = e["dram_bytes_read"] + e["dram_bytes_write"]
        e["time_us"] = round(e["time_us"], 3)
        e["source"] = f"{txt} (ncu --set full of profiles/prof_frame.py, config C3, {note})"
    bins = [e for k, e in fam.items() if k in ("onesweep_kernel", "scan_emit_kernel", "tile_ranges_kernel")]
    fam["binning_total"] = {"kernel": " + ".join(f"{k.replace('_kernel', '')} x{e['launches']}" for k, e in fam.items()
                                                   if k in ("onesweep_kernel", "scan_emit_kernel", "tile_ranges_kernel")),
                            "dram_bytes": sum(e["dram_bytes"] for e in bins), "time_us": round(sum(e["time_us"] for e in bins), 3),
                            "source": txt}
    json.dump(fam, open(js, "w"), indent=1)
    for k, e in fam.items():
        if "launches" not in e:
            continue
        print(f"{k:32s} n={e['launches']} {e['time_us']:8.1f} us  {e['dram_bytes']/1e6:8.1f} MB  {e['warp_inst']/1e6:7.1f} Minst")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "the final code")
