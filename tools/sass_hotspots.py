"""SASS-level hot spots of every kernel in an `ncu --set full --import-source on` report: executed warp instructions by
opcode, stall samples by reason, and the hottest basic blocks (runs of SASS lines with equal execution count).
    python tools/sass_hotspots.py <report.ncu-rep> <out.txt> "<comment line>"          (runs here, no GPU)"""
import collections
import csv
import subprocess
import sys

rep, out_path, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["data"].append(r)
with open(out_path, "w") as f:
    f.write(f"# {note}\n# per kernel: executed warp instructions by opcode, stall samples by reason, and the hottest basic blocks "
            f"(runs of SASS lines with equal execution count)\n")
    seen = set()
    for k in kernels:
        key = (k["name"], len(k["data"]), tuple(k["data"][0]) if k["data"] else ())
        if key in seen:  # ncu prints the first kernel of a multi-kernel report twice
            continue
        seen.add(key)
        ix = {h: i for i, h in enumerate(k["hdr"])}
        stall = [h for h in k["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
        def num(r, h):
            try:
                return int(float(r[ix[h]] or 0))
            except ValueError:
                return 0
        ops, st = collections.Counter(), collections.Counter()
        tot_i = tot_s = 0
        for r in k["data"]:
            e = num(r, "Instructions Executed")
            tok = r[ix["Source"]].split()
            op = next((t for t in tok if not t.startswith("@")), "?").split(".")[0]
            ops[op] += e
            tot_i += e
            tot_s += num(r, "# Samples")
            for h in stall:
                st[h[6:]] += num(r, h)
        f.write(f"\n== {k['name']}\n   executed warp instructions {tot_i}, stall samples {tot_s}\n")
        f.write("   opcodes (% of executed): " + ", ".join(f"{o} {100.0 * c / max(tot_i, 1):.1f}" for o, c in ops.most_common(16)) + "\n")
        f.write("   stall samples: " + ", ".join(f"{o} {c}" for o, c in st.most_common(10)) + "\n")
        blocks, i = [], 0
        d = k["data"]
        while i < len(d):
            e = num(d[i], "Instructions Executed")
            j = i
            while j + 1 < len(d) and num(d[j + 1], "Instructions Executed") == e:
                j += 1
            if e > 0:
                bo = collections.Counter()
                for r in d[i:j + 1]:
                    tok = r[ix["Source"]].split()
                    bo[next((t for t in tok if not t.startswith("@")), "?").split(".")[0]] += 1
                blocks.append(((j - i + 1) * e, j - i + 1, e, sum(num(r, "# Samples") for r in d[i:j + 1]), bo))
            i = j + 1
        for w, n, e, s, bo in sorted(blocks, key=lambda b: -b[0])[:8]:
            f.write(f"   block: {n:4d} instr x {e:8d} executions = {100.0 * w / max(tot_i, 1):5.1f} % of executed, {s:5d} samples; "
                    + ", ".join(f"{o} {c}" for o, c in bo.most_common(7)) + "\n")
print(open(out_path).read())
