#!/usr/bin/env python
"""Turn gpurun_out ncu outputs into the tracked summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches_v1.md
    python tools/ncu_summary.py full gpurun_out/prof_r1_v1.ncu-rep profiles/r1_full_v1.md
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[h], rows[h + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        name = r[ki].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list: {src}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over one "
                f"hot-path step (cold-cache, serialised: compare shares).\n\n"
                f"total {tot / 1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches\n\n"
                "| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.3f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary: {src}\n\n(`--clock-control none --import-source on`; one row group per "
                "captured launch)\n")
        for r in data:
            f.write(f"\n## `{r[ki].split('(')[0]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in KEEP:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"| {m} | {r[i]} | {units[i]} |\n")
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
