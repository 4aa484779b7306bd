#!/usr/bin/env python3
"""Turns an .ncu-rep (ncu --set full, one launch per kernel) into the per-kernel summary kept under profiles/ and, with
--traffic N_TRACKS, into profiles/traffic.json (DRAM bytes per track of every kernel: what bench.py reports as
roofline.traffic).  usage: ncu_summary.py report.ncu-rep out.csv "header comment" [--traffic N out.json]"""
import csv
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
rep, out_csv, comment = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in data if not r[col["Kernel Name"]].startswith("k_fetch_desc")]   # descriptor fetches: microseconds, not part of the summary
names = [r[col["Kernel Name"]] for r in data]
short = [n.split("(")[0].replace("void ", "") for n in names]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
with open(out_csv, "w", newline="") as fh:
    fh.write("# " + comment + "\n")
    w = csv.writer(fh)
    w.writerow(["metric", "unit"] + short)
    w.writerow(["Block Size", ""] + [r[col["Block Size"]] for r in data])
    w.writerow(["Grid Size", ""] + [r[col["Grid Size"]] for r in data])
    for m in METRICS + stall:
        if m in col:
            w.writerow([m, units[col[m]]] + [r[col[m]] for r in data])
if "--traffic" in sys.argv:
    k = sys.argv.index("--traffic")
    n_tracks, out_json = int(sys.argv[k + 1]), sys.argv[k + 2]

    def gb(r, m):
        v, u = float(r[col[m]]), units[col[m]]
        return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}[u]
    # units are per column in ncu's csv (one unit row), so all kernels share it
    rd = {s: gb(r, "dram__bytes_read.sum") for s, r in zip(short, data)}
    wr = {s: gb(r, "dram__bytes_write.sum") for s, r in zip(short, data)}
    try:
        old = json.load(open(out_json))
    except Exception:  # noqa: BLE001
        old = {}
    # a kernel this capture did not reach (ncu -c N stops after N launches) keeps the figures of the previous capture
    kept = []
    for okey, oval in old.items():
        if okey.startswith("per_kernel_gbytes_%d_tracks" % n_tracks):
            for name, v in oval["read"].items():
                if name not in rd:
                    rd[name] = v; wr[name] = oval["write"].get(name, 0.0); kept.append(name)
    if kept:
        comment += " (" + ", ".join(kept) + ": unchanged kernels, figures of the previous capture)"
    tj = {"source": comment + ": dram__bytes_read.sum + dram__bytes_write.sum of the first launch of each kernel, divided by %d tracks (%s)" % (n_tracks, out_csv),
          "k_hybrid_synth_dram_bytes_per_track": (rd.get("k_hybrid_synth", 0) + wr.get("k_hybrid_synth", 0)) * 1e9 / n_tracks,
          "all_kernels_dram_bytes_per_track": (sum(rd.values()) + sum(wr.values())) * 1e9 / n_tracks,
          "per_kernel_gbytes_%d_tracks" % n_tracks: {"read": rd, "write": wr}}
    # warp-instructions per track of every kernel (issue roofline of bench.py)
    inst = {s_: float(r[col["smsp__inst_executed.sum"]]) / n_tracks for s_, r in zip(short, data) if "smsp__inst_executed.sum" in col}
    tj["warp_instructions_per_track"] = inst
    for key in ("k_gf_scan_dram_bytes_per_track", "k_gf_scan_source"):
        if key in old:
            tj[key] = old[key]
    json.dump(tj, open(out_json, "w"), indent=1)
print("wrote", out_csv)
