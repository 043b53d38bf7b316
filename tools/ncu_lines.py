"""tuning aid: per-source-line instruction counts and stall samples from an .ncu-rep
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in want or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(vals[i] or 0) > 0.05):
        print(h, units[i], vals[i])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur = None; hd = None; agg = []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hd = r; continue
    if hd and len(r) == len(hd) and r[0] != "":
        d = dict(zip(hd, r))
        try: agg.append((cur, int(r[0]), r[1].strip()[:84], int(d["Instructions Executed"]), int(d["# Samples"]), int(d.get("stall_long_sb", 0) or 0)))
        except Exception: pass
ti = sum(a[3] for a in agg); ts = sum(a[4] for a in agg)
print("total inst", ti, "samples", ts)
for key, name in ((3, "instructions"), (4, "samples")):
    print("---- by", name)
    for a in sorted(agg, key=lambda a: -a[key])[:top]:
        print("%-22s %4d inst %5.1f%% samp %5.1f%% longsb %5.1f%% | %s" % (a[0], a[1], 100 * a[3] / ti, 100 * a[4] / ts, 100 * a[5] / ts, a[2]))
