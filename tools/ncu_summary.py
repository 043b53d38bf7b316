"""Summarise one kernel of an .ncu-rep (ncu --set full) as markdown + profiles/traffic.json.
usage: python tools/ncu_summary.py report.ncu-rep "<command line used>" out.md [traffic.json [workload]]
(traffic.json maps a bench workload to the DRAM bytes of one launch of its search kernel; entries of other workloads are kept)"""
import csv, io, json, subprocess, sys

rep, cmd, out = sys.argv[1], sys.argv[2], sys.argv[3]
traffic = sys.argv[4] if len(sys.argv) > 4 else None
workload = sys.argv[5] if len(sys.argv) > 5 else "c2_1Mx128_f32_l2"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def get(name):
    i = col.get(name)
    return (None, None) if i is None else (vals[i], units[i])


def bytes_of(name):
    v, u = get(name)
    return float(v) * SCALE.get(u, 1.0)


want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
kernel = vals[col["Kernel Name"]]
lines = [f"# ncu --set full --clock-control none, one launch of `{kernel}`", "", f"command: `{cmd}`", "", "| metric | value |", "|---|---|"]
for w in want:
    v, u = get(w)
    if v is not None:
        lines.append(f"| {w} | {v} {u} |")
total = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
lines.append(f"| dram traffic per launch (read+write) | {total / 1e9:.3f} GB |")
lines += ["", "warp stall reasons (average warps stalled per issue-active cycle):", "", "| reason | value |", "|---|---|"]
st = []
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        st.append((float(vals[col[h]] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
for v, n in sorted(st, reverse=True)[:10]:
    lines.append(f"| {n} | {v:.2f} |")
lines += ["", "Tensor pipe: not used by design (HBM-gather / integer path; see DESIGN.md §5)."]
open(out, "w").write("\n".join(lines) + "\n")
if traffic:
    try:
        t = json.load(open(traffic))
    except Exception:
        t = {}
    if "search_kernel_dram_bytes_per_launch" in t:  # round-1 layout: a single (C2) entry
        t = {t.get("workload", "c2_1Mx128_f32_l2"): t}
    t[workload] = {"search_kernel_dram_bytes_per_launch": total, "source": f"ncu --set full, {out}", "kernel": kernel}
    json.dump(t, open(traffic, "w"), indent=1)
print(out, "dram GB", total / 1e9)
