"""Summarise ncu outputs into profiles/ (run in the authoring container; ncu reads .ncu-rep without a GPU).

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/<name>.md
  python scripts/summarize_ncu.py full     gpurun_out/prof.ncu-rep  profiles/<name>.md
"""
import collections, csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__inst_executed.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("gsb::<unnamed>::", "gsb::")
        agg.setdefault(name[-70:], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        if sum(v) / tot < 0.002:
            continue
        out.append(f"| `{n}` | {len(v)} | {sum(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
    out.append(f"\ntotal {tot / 1e3:.1f} us over {sum(len(v) for v in agg.values())} launches "
               "(per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes)")
    open(dst, "a").write("\n".join(out) + "\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    out = []
    tr = {}
    for r in rows[2:]:
        name = r[ki].split("(")[0][-40:]
        out.append(f"\n#### `{name}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
        out.append("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"| {k} | {r[i]} | {units[i]} |")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
            mul = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}
            tr[name] = rd * mul[units[hdr.index("dram__bytes_read.sum")]] + wr * mul[units[hdr.index("dram__bytes_write.sum")]]
        except Exception:
            pass
    open(dst, "a").write("\n".join(out) + "\n")
    print(json.dumps(tr))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
