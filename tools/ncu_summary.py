"""Summarise an ncu report for profiles/: python tools/ncu_summary.py <file.ncu-rep> "<title>" [kernel-regex] > profiles/x.md
One table per kernel (last launch of each name), metric names as ncu prints them."""
import csv, io, re, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
col = {}
for i, h in enumerate(hdr):
    for w in WANT:
        if h == w or h.endswith("." + w):
            col.setdefault(w, i)
ki, bi, gi = hdr.index("Kernel Name"), hdr.index("Block Size"), hdr.index("Grid Size")
last = {}
for r in rows[2:]:
    if len(r) == len(hdr) and (pat is None or pat.search(r[ki])):
        last[r[ki]] = r
print(f"# {title}\n\n`ncu --set full --clock-control none --import-source on` (B200). Values per launch (last launch of each kernel).\n")
for name, r in last.items():
    print(f"## `{name}`  grid {r[gi]} block {r[bi]}\n\n| metric | value | unit |\n|---|---|---|")
    for w in WANT:
        if w in col and r[col[w]] != "":
            print(f"| {w} | {r[col[w]]} | {units[col[w]]} |")
    print()
