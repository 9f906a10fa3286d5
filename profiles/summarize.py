#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r01_v6.csv profiles/r01_v6_launches.md
    python profiles/summarize.py full gpurun_out/prof_r01_v6.ncu-rep profiles/r01_v6_cappi3d_fast.md [traffic-key]

`launches` aggregates the `--metrics gpu__time_duration.sum` CSV per kernel (count, total, mean, share);
`full` reads one `ncu --set full` report (needs the `ncu` CLI) and writes the headline metrics, the
executed-instruction mix by opcode and the hottest SASS lines; with a traffic key it also records the
kernel's DRAM bytes per launch in profiles/traffic.json (what bench.py reports as roofline.traffic).
"""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

HEADLINE = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1)


def launches(src, dst):
    rows = []
    with open(src) as fh:
        text = "".join(l for l in fh if l.startswith('"'))
    for r in csv.DictReader(io.StringIO(text)):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    agg = {}
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    ours = sum(a[1] for n, a in agg.items() if "pcg::" in n)
    with open(dst, "w") as out:
        out.write("# ncu launch list — `%s`\n\n" % os.path.basename(src))
        out.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over one `bench.py` run "
                  "(cold-cache, serialised launches: shares matter, not absolutes).\n\n")
        out.write("| kernel | launches | total µs | mean µs | share of all GPU time | grid | block |\n|---|---|---|---|---|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write("| `%s` | %d | %.1f | %.2f | %.1f%% | %s | %s |\n" % (
                name[:110], a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / total, a[2], a[3]))
        out.write("\nTotal GPU time %.1f µs over %d launches; kernels of this library (`pcg::`): %.1f%% "
                  "(the rest is bench.py's L2 flush write and torch's H2D helpers).\n" % (total / 1e3, len(rows), 100 * ours / total))
    print("wrote", dst)


def ncu_csv(rep, page, extra=()):
    cmd = ["ncu", "-i", rep, "--page", page, "--csv"] + list(extra)
    text = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(text)))


def full(rep, dst, traffic_key=None):
    raw = ncu_csv(rep, "raw")
    hdr, units, row = raw[0], raw[1], raw[2]
    metric = {h: (row[i], units[i]) for i, h in enumerate(hdr)}
    kernel = metric.get("Kernel Name", ("?", ""))[0]
    sass = ncu_csv(rep, "source", ["--print-source", "sass"])
    shdr = sass[1]
    i_src, i_exec, i_smp = shdr.index("Source"), shdr.index("Instructions Executed"), shdr.index("# Samples")
    lines, byop = [], {}
    for r in sass[2:]:
        try:
            n, s = int(r[i_exec]), int(r[i_smp])
        except (ValueError, IndexError):
            continue
        toks = r[i_src].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        a = byop.setdefault(op, [0, 0])
        a[0] += n
        a[1] += s
        lines.append((s, n, r[i_src].strip()))
    tot = sum(a[0] for a in byop.values()) or 1
    tot_s = sum(a[1] for a in byop.values()) or 1
    rd = to_bytes(*metric["dram__bytes_read.sum"])
    wr = to_bytes(*metric["dram__bytes_write.sum"])
    with open(dst, "w") as out:
        out.write("# ncu --set full — `%s`\n\nKernel: `%s`\n\n" % (os.path.basename(rep), kernel))
        out.write("| metric | value | unit |\n|---|---|---|\n")
        for k in HEADLINE:
            if k in metric:
                out.write("| %s | %s | %s |\n" % (k, metric[k][0], metric[k][1]))
        out.write("\nDRAM traffic per launch: read %.1f MB + write %.1f MB = **%.1f MB**.\n" % (rd / 1e6, wr / 1e6, (rd + wr) / 1e6))
        out.write("\n## Executed warp instructions by opcode (total %d)\n\n| opcode | executed | share | stall samples | sample share |\n|---|---|---|---|---|\n" % tot)
        for op, a in sorted(byop.items(), key=lambda kv: -kv[1][0])[:28]:
            out.write("| %s | %d | %.1f%% | %d | %.1f%% |\n" % (op, a[0], 100 * a[0] / tot, a[1], 100 * a[1] / tot_s))
        out.write("\n## Hottest SASS lines by stall samples\n\n| samples | executed | SASS |\n|---|---|---|\n")
        for s, n, text in sorted(lines, key=lambda t: -t[0])[:25]:
            out.write("| %d | %d | `%s` |\n" % (s, n, text))
    print("wrote", dst)
    if traffic_key:
        tpath = os.path.join(HERE, "traffic.json")
        data = {}
        if os.path.exists(tpath):
            data = json.load(open(tpath))
        data[traffic_key] = {"bytes": rd + wr, "read": rd, "write": wr, "kernel": kernel, "report": os.path.basename(rep)}
        json.dump(data, open(tpath, "w"), indent=1, sort_keys=True)
        print("updated", tpath)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif mode == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        raise SystemExit(__doc__)
