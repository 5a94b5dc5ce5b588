#!/usr/bin/env python3
"""Turn an ncu --set full report into the committed evidence: a per-kernel summary CSV and profiles/ncu_traffic.json (DRAM
bytes per launch of the correlation kernels, keyed by the digest of the library sources the capture was taken on -- bench.py
only reports `roofline.traffic` when that digest matches the sources it runs).

    python tools/ncu_summary.py gpurun_out/<tag>_prof.ncu-rep <digest> profiles/r2/<name>.csv
"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, digest, out_csv = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
idx = [(k, hdr.index(k)) for k in KEEP if k in hdr]
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([k for k, _ in idx])
    w.writerow([units[i] for _, i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for _, i in idx])
iN, iR, iW = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {}
for r in rows[2:]:
    for key in ("corr_bwd_tc_kernel", "corr_fwd_tc_kernel", "corr_tc_split_kernel"):
        if key in r[iN] and key not in traffic:
            traffic[key] = int(float(r[iR]) * scale[units[iR]] + float(r[iW]) * scale[units[iW]])
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
table = json.load(open(path)) if os.path.isfile(path) else {}
traffic["source"] = os.path.relpath(out_csv, ROOT) + " (ncu --set full --clock-control none, tools/prof_ops.py all 1, cfg2)"
table = {digest: traffic}          # only the capture of the shipped sources is kept
json.dump(table, open(path, "w"), indent=1)
print(json.dumps(table, indent=1))
