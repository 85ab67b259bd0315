"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py <csv> "<title>" """
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows:
    v = float(r[14].replace(",", "")); u = r[13]
    us = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
    tot[r[4]] += us; cnt[r[4]] += 1
allus = sum(tot.values()); ours = sum(v for k, v in tot.items() if "b200::" in k)
print(f"# {sys.argv[2]}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400`, run right after the same command exited 0 without ncu.\n"
      "Per-launch times are cold-cache and serialised; compare SHARES.  `at::…` kernels are bench.py's own input generation and\n"
      "L2 flush (torch), not the product path; the seven `b200::reduce_*` kernels are one configs[1] step.\n")
print(f"Share of `b200::` kernels in the captured GPU time: {100 * ours / allus:.1f} % (the rest is input generation / L2 flush).\n")
print("| kernel | launches | total us | share of all | share of b200:: |\n|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    sh = f"{100 * v / ours:.1f}%" if "b200::" in k else "—"
    print(f"| `{k[:100]}` | {cnt[k]} | {v:.1f} | {100 * v / allus:.1f}% | {sh} |")
