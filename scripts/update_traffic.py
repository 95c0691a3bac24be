"""profiles/traffic.json <- the steady-state DRAM traffic per step_obs_kernel launch from an ncu capture of THIS round:
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none
       -k regex:step_obs -s 16 -c 16 --csv --log-file <csv> python scripts/prof_step.py
usage: update_traffic.py <csv> <key, e.g. step_obs_v17_n65536> [<csv> <key> ...]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "traffic.json")
out = {}
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or os.environ.get("TA_COMMIT", "unknown")
detail = {}
for f, key in zip(sys.argv[1::2], sys.argv[2::2]):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    h = rows[0]; iM, iV, iID = h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = {}
    for r in rows[1:]:
        per.setdefault(r[iID], {})[r[iM]] = float(r[iV].replace(",", ""))
    n = len(per)
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in per.values()) / n
    wr = sum(d.get("dram__bytes_write.sum", 0) for d in per.values()) / n
    us = sum(d.get("gpu__time_duration.sum", 0) for d in per.values()) / n / 1e3
    out[key] = int(rd + wr)
    detail[key] = {"launches": n, "dram_read_bytes": int(rd), "dram_write_bytes": int(wr), "us_under_ncu": round(us, 2), "csv": os.path.basename(f)}
out["source"] = (f"ncu --cache-control none --clock-control none -k regex:step_obs -s 16 -c 16 on scripts/prof_step.py (8 rotating batches), "
                 f"mean over the captured launches; round 2, taken at commit {head}; committed capture, not measured inside bench.py")
out["detail"] = detail
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
