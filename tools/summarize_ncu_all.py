"""Per-kernel table from an `ncu --set full` report that holds one launch of every kernel of the path.
    python tools/summarize_ncu_all.py <report.ncu-rep> <out.txt> <clips>"""
import csv
import subprocess
import sys

rep, out_path, clips = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}


def val(r, name, default=0.0):
    try:
        return float(r[ci[name]].replace(",", ""))
    except Exception:
        return default


def to(v, unit, scale):
    return v * scale.get(unit, 1.0)


lines = []
for r in rows[2:]:
    name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "").replace("avs::", "")[:34]
    dur_us = to(val(r, "gpu__time_duration.sum"), units[ci["gpu__time_duration.sum"]], {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6})
    rd = to(val(r, "dram__bytes_read.sum"), units[ci["dram__bytes_read.sum"]], {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3})
    wr = to(val(r, "dram__bytes_write.sum"), units[ci["dram__bytes_write.sum"]], {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3})
    gbs = (rd + wr) / 1e3 / (dur_us / 1e6) if dur_us else 0.0
    lines.append((name, r[ci["launch__grid_size"]], r[ci["launch__block_size"]], dur_us, rd, wr, gbs,
                  val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                  val(r, "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
                  val(r, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                  val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                  val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                  val(r, "launch__registers_per_thread")))
with open(out_path, "w") as f:
    f.write(f"# ncu --set full --clock-control none, one launch of every kernel of the path, {clips} clips (bf16 STCNN)\n")
    f.write("# tools/prof_sweep.py; times are single cold launches under the profiler: use for RATIOS and traffic\n")
    f.write(f"{'kernel':34s} {'grid':>7s} {'blk':>5s} {'us':>9s} {'us/clip':>8s} {'rdMB':>8s} {'wrMB':>8s} {'GB/s':>7s} {'dram%':>6s} "
            f"{'tens%':>6s} {'tcsm%':>6s} {'sm%':>6s} {'warps%':>6s} {'regs':>5s}\n")
    for (n, g, b, d, rd, wr, gbs, dp, tp, ts, sp, wa, rg) in lines:
        f.write(f"{n:34s} {g:>7s} {b:>5s} {d:9.1f} {d / clips:8.2f} {rd:8.2f} {wr:8.2f} {gbs:7.0f} {dp:6.1f} {tp:6.1f} {ts:6.1f} "
                f"{sp:6.1f} {wa:6.1f} {rg:5.0f}\n")
    f.write("# dram% = gpu__dram_throughput pct of peak; tens% = sm__pipe_tensor_cycles_active_realtime pct of peak;\n"
            "# tcsm% = l1tex__data_pipe_tc_wavefronts_mem_shared pct of peak (tensor-core operand fetch from shared memory)\n")
print(open(out_path).read())
