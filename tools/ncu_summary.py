"""Turn an .ncu-rep (ncu --set full) into the markdown table used under profiles/ (development aid).

    python tools/ncu_summary.py <report.ncu-rep> [kernel regex]
One table per distinct kernel name (first launch of each), plus the stall reasons per issued instruction.
"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"]
seen = set()
for r in data:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "")
    short = re.sub(r"\(.*", "", name)
    if (pat and not pat.search(name)) or short in seen:
        continue
    seen.add(short)
    u = dict(zip(hdr, units))
    print(f"### `{short}`\n\n| metric | value |\n|---|---|")
    for k in KEYS:
        if k in d and d[k] != "":
            print(f"| {k} | {d[k]} {u.get(k, '')} |")
    stalls = sorted(((k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), float(d[k]))
                     for k in hdr if "smsp__average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio")
                     and "not_issued" not in k and d[k] not in ("", "n/a")), key=lambda kv: -kv[1])[:6]
    print("\nWarp stall reasons (cycles per issued instruction): " + ", ".join(f"{k} {v:.2f}" for k, v in stalls))
    try:
        t = float(d["gpu__time_duration.sum"]); rd = float(d["dram__bytes_read.sum"]); wr = float(d["dram__bytes_write.sum"])
        tu, bu = u["gpu__time_duration.sum"], u["dram__bytes_read.sum"]
        ts = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-9)
        sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rd * sc.get(bu, 1.0) + wr * sc.get(u["dram__bytes_write.sum"], 1.0)
        print(f"\nDRAM traffic {tot / 1e6:.1f} MB in {ts * 1e6:.1f} us = {tot / ts / 1e9:.0f} GB/s\n")
    except (KeyError, ValueError):
        print()
