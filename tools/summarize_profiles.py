"""Turn the raw ncu outputs under gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_profiles.py <round-tag> <launches.csv> <full.ncu-rep> [clips-in-full-capture]
"""
import collections
import csv
import re
import subprocess
import sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
clips = int(sys.argv[4]) if len(sys.argv) > 4 else 8

rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
ci = {h: i for i, h in enumerate(rows[0])}
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[ci["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(r[ci["Metric Unit"]], 1.0)
    a = agg.setdefault(r[ci["Kernel Name"]].split("(")[0][:70], [0, 0.0, r[ci["Grid Size"]], r[ci["Block Size"]]])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(f"profiles/{tag}_launches.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write(f"# source: {launches}; command: python bench.py --steps 2 --warmup 3 --clips-per-gpu 256 --cpu-clips 0 --no-e2e --no-configs (tools/ncu_round.sh)\n")
    f.write(f"{'kernel':72s} {'n':>5s} {'total ms':>10s} {'share':>7s} {'avg us':>10s}  grid / block\n")
    for k, (n, t, g, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k:72s} {n:5d} {t / 1e3:10.3f} {t / tot:7.1%} {t / n:10.1f}  {g} / {b}\n")
print(open(f"profiles/{tag}_launches.txt").read())

out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
hdr, units = r[0], r[1]
keep = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
GFLOP = {"0": 1.8, "1": 28.8, "2": 7.46}  # algorithmic GFLOP per clip of the three layers
with open(f"profiles/{tag}_conv2_ncu_full.txt", "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k 'regex:conv_umma|conv_l2_fused' -s 3 -c 3 python tools/prof_conv.py\n")
    f.write(f"# the three conv launches of one STCNN forward (conv_umma_kernel<0> = layer 1, conv_l2_fused_kernel = layer 2, conv_umma_kernel<2> = layer 3), {clips} clips per launch (PROF_CLIPS), bf16; "
            f"layer 2 is the roofline kernel of bench.py\n")
    for vals in r[2:]:
        d = {}
        f.write("\n")
        for h, u, v in zip(hdr, units, vals):
            if h in keep:
                f.write(f"{h:95s} {v} {u}\n")
                d[h] = (v, u)
        try:
            rd, wr = float(d["dram__bytes_read.sum"][0].replace(",", "")), float(d["dram__bytes_write.sum"][0].replace(",", ""))
            kind = "1" if "fused" in d["Kernel Name"][0] else re.search(r"<(?:\(int\))?(\d)>", d["Kernel Name"][0]).group(1)
            f.write(f"# derived: DRAM traffic per launch = {rd + wr:.1f} {d['dram__bytes_read.sum'][1]} = {(rd + wr) / clips:.2f} per clip; "
                    f"algorithmic {GFLOP[kind]} GFLOP/clip\n")
        except Exception as e:
            f.write(f"# derived: n/a ({e})\n")
print(open(f"profiles/{tag}_conv2_ncu_full.txt").read())
