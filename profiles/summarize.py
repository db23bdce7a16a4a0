"""Turns ncu outputs brought back in gpurun_out/ into the compact summaries kept under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv  > profiles/X_launches.md
    python profiles/summarize.py report   gpurun_out/prof_X.ncu-rep  > profiles/X_top_kernels.md
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    mi = hdr.index("Metric Name")
    agg, order, total, nl = {}, [], 0.0, 0
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("<unnamed>::", "")[:90]
        v = float(r[vi].replace(",", ""))
        if name not in agg:
            agg[name] = [0, 0.0, 0.0]
            order.append(name)
        if r[mi] == "gpu__time_duration.sum":
            v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)   # -> us
            agg[name][0] += 1
            agg[name][1] += v
            total += v
            nl += 1
        elif r[mi].startswith("dram__bytes"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
            agg[name][2] += v * scale
    print("| kernel | launches | total us | share | avg us | avg DRAM MB (read + write) |\n|---|---:|---:|---:|---:|---:|")
    for name in sorted(order, key=lambda n: -agg[n][1]):
        c, t, b = agg[name]
        print(f"| `{name}` | {c} | {t:.1f} | {100 * t / total:.1f}% | {t / c:.1f} | {b / c / 1e6:.1f} |")
    print(f"\ntotal {total:.1f} us over {nl} launches (ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --clock-control none: cold-cache, serialised)")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("###", r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", ""), "\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
