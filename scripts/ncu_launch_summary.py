"""Aggregate an `ncu --csv --metrics ...` launch list by kernel: launches, total us, share of the step,
time-weighted tensor-pipe and DRAM utilisation, DRAM bytes.   usage: ncu_launch_summary.py list.csv [top_n]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[0]; iN, iM, iV, iID = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
k = collections.OrderedDict()
for r in rows[1:]:
    k.setdefault(r[iID], {'name': r[iN]})[r[iM]] = float(r[iV].replace(',', ''))
T = 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'; D = 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'
agg = collections.OrderedDict(); tot = 0.0
for d in k.values():
    t = d['gpu__time_duration.sum'] / 1e3
    a = agg.setdefault(d['name'][:72], [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += t * d.get(T, 0); a[3] += t * d.get(D, 0); a[4] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    tot += t
print(f"{len(k)} launches, {tot:.0f} us of kernel time (ncu: serialised, cold caches; shares are what to read)")
print(f"{'kernel':72s} {'n':>3s} {'us':>8s} {'share':>6s} {'tensor%':>7s} {'dram%':>6s} {'DRAM MB':>8s}")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:topn]:
    print(f"{n:72s} {a[0]:3d} {a[1]:8.1f} {a[1] / tot * 100:5.1f}% {a[2] / a[1]:7.1f} {a[3] / a[1]:6.1f} {a[4] / 1e6:8.1f}")
