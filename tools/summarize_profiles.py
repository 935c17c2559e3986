"""Turns gpurun_out/ ncu artefacts into the tracked text summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_r1.csv  profiles/r1_launches.md
    python tools/summarize_profiles.py full     gpurun_out/prof_c3_r1.ncu-rep profiles/r1_c3_full.md
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_xu.sum", "smsp__cycles_active.avg",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v = v / 1e3 if u.startswith("n") else v * 1e3 if u.startswith("m") else v
        agg.setdefault(r[ki], []).append((v, r[gi], r[bi]))
    total = sum(v for vs in agg.values() for v, _, _ in vs)
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised:\n"
                "# compare SHARES, not absolutes) — source: %s\n\n" % src)
        f.write("| kernel | launches | mean us | total us | share | grid x block |\n|---|---|---|---|---|---|\n")
        for k, vs in agg.items():
            t = sum(v for v, _, _ in vs)
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% | %s x %s |\n" % (
                k[:110], len(vs), t / len(vs), t, 100 * t / total, vs[0][1], vs[0][2]))
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none — source: %s\n" % src)
        for r in rows[2:]:
            f.write("\n## %s\n\n| metric | value | unit |\n|---|---|---|\n" % r[idx["Kernel Name"]][:160])
            for m in METRICS:
                if m in idx:
                    f.write("| %s | %s | %s |\n" % (m, r[idx[m]], units[idx[m]]))
            try:
                rd = float(r[idx["dram__bytes_read.sum"]])
                wr = float(r[idx["dram__bytes_write.sum"]])
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                tot = rd * scale[units[idx["dram__bytes_read.sum"]]] + wr * scale[units[idx["dram__bytes_write.sum"]]]
                dur = float(r[idx["gpu__time_duration.sum"]])
                du = units[idx["gpu__time_duration.sum"]]
                sec = dur * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "s": 1.0, "second": 1.0}[du]
                f.write("| **dram traffic (read+write)** | %.0f | byte |\n| **dram GB/s under ncu** | %.0f | GB/s |\n" % (tot, tot / sec / 1e9))
            except Exception as ex:
                f.write("| traffic | n/a (%s) | |\n" % ex)
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
