"""Per-kernel counts of the SASS opcodes that prove what the shipped library runs on: UTCHMMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk), UTMALDG (TMA tensor loads), SYNCS (mbarrier), plus
HMMA / FFMA for contrast.  Runs here (no GPU): python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "alignment-between-speech-and-visual-mouth-movements_b200", "libavsync_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "FFMA", "DFMA", "RED", "ATOM", "SHFL", "STG", "LDG"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, total = collections.OrderedDict(), collections.Counter()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        op = m.group(1)
        counts[cur]["_all"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[cur][o] += 1
                total[o] += 1
h = subprocess.run(["sha256sum", LIB], capture_output=True, text=True).stdout.split()[0][:16]
print(f"# {os.path.basename(LIB)} sha256 {h}  (cuobjdump -sass; arch list: " +
      ",".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass)))) + ")")
print(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{o:>7s}" for o in OPS))
for k, c in counts.items():
    print(f"{k[:78]:78s} {c['_all']:7d} " + " ".join(f"{c[o]:7d}" for o in OPS))
print(f"{'TOTAL':78s} {sum(c['_all'] for c in counts.values()):7d} " + " ".join(f"{total[o]:7d}" for o in OPS))
