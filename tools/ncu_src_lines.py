"""Per-source-line instruction and stall-sample totals from an ncu report:
   ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:K > x.csv; python tools/ncu_src_lines.py x.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
out = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0].strip().isdigit():
        # the source text may contain commas that split it over several fields: numeric tail starts after the two '-' fields
        k = next(i for i in range(2, len(r)) if r[i] == '-' and r[i + 1] == '-')
        samples, inst = int(r[k + 2]), int(r[k + 5])
        out.append((inst, samples, cur, int(r[0]), ','.join(r[1:k])[:110]))
ti = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print(f"total warp instructions {ti}, stall samples {ts}")
for inst, samples, f, l, src in sorted(out, reverse=True)[:top]:
    print(f"{100*inst/ti:5.1f}% inst {100*samples/ts:5.1f}% smp  {f}:{l}  {src.strip()}")
