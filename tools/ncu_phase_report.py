#!/usr/bin/env python
"""Per-phase instruction / stall breakdown of topk_kernel from an ncu report:  python tools/ncu_phase_report.py report.ncu-rep n_users"""
import collections, csv, subprocess, sys
rep, n_users = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct", "launch__grid_size", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct"]
for h, u, v in zip(rows[0], rows[1], rows[2]):
    if h in want:
        print(f"{h:90s} {u:8s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
iS, iI = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
per, cur, curfile = collections.OrderedDict(), None, None
for r in rows:
    if r and r[0] == "File Path":
        curfile = r[1].split("/")[-1]
        continue
    if len(r) <= iI or r[0] == "Line No":
        continue
    if r[0]:
        cur = (curfile, int(r[0]), r[1].strip()[:100])
        per.setdefault(cur, [0, 0])
    elif cur:
        try:
            per[cur][0] += int(r[iS]); per[cur][1] += int(r[iI])
        except ValueError:
            pass
ti, ts = sum(v[1] for v in per.values()), sum(v[0] for v in per.values())
print(f"warp instructions per user: {ti / n_users:.0f}; stall samples {ts}")
for (f, ln, text), (s, i) in sorted(per.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{f}:{ln:<5d} stall {100 * s / ts:5.1f}%  inst {100 * i / ti:5.1f}%  {text}")
