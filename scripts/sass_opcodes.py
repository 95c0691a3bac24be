"""Per-kernel counts of the SASS opcodes that prove the sm_100a features used (tcgen05 MMA = UTCHMMA, TMEM loads = LDTM,
MMA barriers = UTCBAR, TMA bulk copies = UBLKCP, packed fp32 FMA = FFMA2, byte permutes = PRMT, griddepcontrol = ACQBULK /
PREEXIT ...).  cuobjdump -sass of the in-tree library; writes profiles/r2_sass_opcodes.txt.  Runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "goal-conditioned-reinforcement-learning-with-environmental-and-policy-priors_b200", "csrc", "libtwoarmy_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "UTMASTG", "FFMA2", "PRMT", "SHFL", "LDG.E.128", "STG.E.128", "STS.128", "LDS.128",
         "ACQBULK", "PREEXIT", "SYNCS", "ATOMG", "RED", "HMMA", "LDGSTS"]
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[kern][w] += 1
lines = [f"cuobjdump -sass {os.path.relpath(so, ROOT)} (nvcc -gencode arch=compute_100a,code=sm_100a); instructions per kernel and the watched opcodes",
         f"git HEAD {subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()}", ""]
grand = collections.Counter()
for k, c in counts.items():
    grand.update(c)
    lines.append(f"{k[:110]:110s} {total[k]:6d} instr  " + "  ".join(f"{w}={n}" for w, n in c.items()))
lines += ["", "TOTAL  " + "  ".join(f"{w}={n}" for w, n in sorted(grand.items()))]
p = os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt")
open(p, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-3:]))
print("wrote", p, len(counts), "kernels")
