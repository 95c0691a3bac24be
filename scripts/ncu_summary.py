"""Summarise an .ncu-rep: key metrics + per-source-line instruction/stall aggregation.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [top_n] [launch index in the report, default 0]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40; which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
def g(k): return rows[2 + which][h.index(k)] if k in h else None
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum']
for k in keys: print(f"{k:70s} {g(k)}")
for k in h:
    if 'issue_stalled' in k and 'per_issue_active' in k:
        print(f"  stall {k.split('stalled_')[1].split('_per')[0]:22s} {g(k)}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; agg = collections.OrderedDict(); hdr = None; inst = 0
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'Kernel Name':
        inst += 1
        if inst > which + 1: break
        agg.clear(); hdr = None
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; iI = hdr.index('Instructions Executed'); iW = hdr.index('Warp Stall Sampling (All Samples)'); continue
    if hdr and r[0].isdigit():
        try:
            k = (cur, int(r[0])); a = agg.get(k, (0, 0, r[1][:100])); agg[k] = (a[0] + int(r[iI]), a[1] + int(r[iW]), r[1][:100])
        except Exception: pass
tot = sum(v[0] for v in agg.values()) or 1; tots = sum(v[1] for v in agg.values()) or 1
print(f"\ntotal warp-inst {tot}  stall samples {tots}")
print("--- by instructions")
for (f, l), (i, w, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{f}:{l:4d} inst={100*i/tot:5.1f}% stall={100*w/tots:5.1f}% | {s}")
print("--- by stall samples")
for (f, l), (i, w, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{f}:{l:4d} inst={100*i/tot:5.1f}% stall={100*w/tots:5.1f}% | {s}")
