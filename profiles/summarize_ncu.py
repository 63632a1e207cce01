"""Dump the judged metrics of an .ncu-rep (read here with `ncu -i ... --page raw --csv`) as a small text file.

    python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...]  > profiles/r1_x.txt
    python profiles/summarize_ncu.py --launches gpurun_out/launches_x.csv          > profiles/r1_launches_x.txt
"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = csv.reader(io.StringIO(out))
    hdr, units = next(rd), next(rd)
    print(f"## {path}")
    for r in rd:
        print(f"### {r[hdr.index('Kernel Name')][:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:88s} {r[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.2:
                    stalls.append((v, h.split("stalled_")[1].split("_per")[0]))
        print("  stall reasons (warps per issue-active cycle, >= 0.2): "
              + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)))


def launches(path):
    lines = [l for l in open(path).read().splitlines() if not l.startswith("==")]
    rd = csv.reader(io.StringIO("\n".join(lines)))
    hdr = next(rd)
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    print(f"## {path}: gpu__time_duration.sum per launch (ncu --clock-control none; cold-cache, serialised)")
    agg = {}
    for r in rd:
        name = re.sub(r"\(.*", "", r[ik])
        if name.startswith("void at::") or "at::native" in name or "distribution_elementwise" in name:
            continue
        us = float(r[iv].replace(",", "")) / 1e3
        agg.setdefault((name, r[ig]), []).append(us)
    tot = sum(sum(v) for v in agg.values())
    for (name, grid), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {name[:52]:52s} grid {grid:16s} n={len(v):5d} mean {sum(v) / len(v):9.2f} us  "
              f"total {sum(v):10.1f} us ({100 * sum(v) / tot:4.1f} %)")


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--launches":
        for p in args[1:]:
            launches(p)
    else:
        for p in args:
            rep(p)
