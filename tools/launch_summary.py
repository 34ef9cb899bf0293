"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py file.csv [top]"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h = r; start = i; break
ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) <= vi: continue
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('cqvad::<unnamed>::', '')[:60]
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    a = agg[name]; a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot / 1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{a[1] / 1e3:10.1f} us {100 * a[1] / tot:5.1f}% {a[0]:6d}  {k}")
