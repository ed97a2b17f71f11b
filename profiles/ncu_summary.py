"""Summarise an ncu report (one row per profiled launch) and derive the bench's traffic table from it:

    python profiles/ncu_summary.py <rep> <out.csv> [--traffic profiles/traffic.json --particles N --note "..."]

<out.csv> is the committed evidence; traffic.json (read by bench.py for `roofline.traffic`) holds, per step kernel, the
mean `dram__bytes_read.sum + dram__bytes_write.sum` per launch and the mean `gpu__time_duration` of the SAME capture,
with the CSV's name, so the number in the bench line can be traced to its rows."""
import csv
import json
import os
import subprocess
import sys

SLOT_OF = {"k_begin_tick": "clear", "k_prepass": "prepass_wall_key", "k_scan_lookback": "scan", "k_place": "place",
           "k_rank_gather": "rank_gather", "k_density": "density", "k_force": "force_integrate",
           "k_dist_pack": "dist_pack", "k_dist_unpack": "dist_unpack", "k_emit": "io_scatter"}

args = sys.argv[1:]
rep, out = args[0], args[1]
opt = dict(zip(args[2::2], args[3::2]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
idx = [hdr.index(w) for w in want if w in hdr]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i][:80] for i in idx])
print(open(out).read())

if "--traffic" in opt:
    def scale(unit):
        return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3}.get(unit, 1.0)
    iname, it, ir, iw = (hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum",
                                                 "dram__bytes_write.sum"))
    acc = {}
    for r in rows[2:]:
        slot = next((v for k, v in SLOT_OF.items() if k in r[iname]), None)
        if slot is None:
            continue
        a = acc.setdefault(slot, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += (float(r[ir]) * scale(units[ir]) + float(r[iw]) * scale(units[iw]))
        a[2] += float(r[it]) * scale(units[it])
    table = {"_source": f"{os.path.basename(out)} (ncu --set full --clock-control none; mean over the profiled launches)"}
    for slot, (n, b, t) in acc.items():
        table[slot] = {"dram_bytes": int(b / n), "ncu_kernel_us": round(t / n, 2), "launches": n,
                       "ncu_csv": "profiles/" + os.path.basename(out), "particles": int(opt.get("--particles", 0)),
                       "note": opt.get("--note", "")}
    json.dump(table, open(opt["--traffic"], "w"), indent=1)
    print(json.dumps(table, indent=1))
