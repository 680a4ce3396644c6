#!/usr/bin/env python
"""Turns ncu artefacts under gpurun_out/ into the small text summaries committed in profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1h.csv profiles/r1h_launches.md
    python profiles/summarize.py full     gpurun_out/prof_r1h.ncu-rep  profiles/r1h_anneal_full.md

`launches`: the per-launch list of `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache,
serialised times: the kernel's SHARE of the step is what must agree with bench.py, not the absolute).
`full`: selected metrics of one `ncu --set full --import-source on` capture + the hottest source lines and the
warp-stall mix (sampled).  Reads the .ncu-rep with the local `ncu -i` (no GPU needed).
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum.per_cycle_elapsed", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    kn, val = h.index("Kernel Name"), h.index("Metric Value")
    per = collections.OrderedDict()
    order = []
    for r in rows[hdr + 1:]:
        if len(r) <= val:
            continue
        name = r[kn].split("(")[0].replace("void ", "")
        ns = float(r[val].replace(",", ""))
        per.setdefault(name, []).append(ns)
        order.append((name, ns))
    tot = sum(ns for _, ns in order)
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over one bench.py run; "
                "times are cold-cache and serialised, so shares (not absolutes) are comparable with bench.py.\n\n")
        f.write("| kernel | launches | total ms | share | mean ms |\n|---|---|---|---|---|\n")
        for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{name}` | {len(v)} | {sum(v) / 1e6:.3f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v) / 1e6:.3f} |\n")
        f.write(f"\ntotal device time in the list: {tot / 1e6:.3f} ms over {len(order)} launches\n")


def full(rep, dst, top=25):
    raw = ncu_csv(rep, "raw")
    h, units = raw[0], raw[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({rep})\n\n")
        for r in raw[2:]:
            d = dict(zip(h, r))
            f.write(f"## `{d['Kernel Name']}`  grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEEP:
                if k in d and d[k] != "":
                    f.write(f"| {k} | {d[k]} | {units[h.index(k)]} |\n")
            f.write("\n")
        src = ncu_csv(rep, "source", ("--print-source", "cuda,sass"))
        hi = [i for i, r in enumerate(src) if r and r[0] == "Line No"]
        files = [i for i, r in enumerate(src) if r and r[0] == "File Path"]
        agg, stalls, tot = [], collections.Counter(), 0
        for s, hidx in enumerate(hi):
            end = files[s + 1] if s + 1 < len(files) else len(src)
            fname = src[files[s]][1].split("/")[-1]
            hh = src[hidx]
            ie, sm = hh.index("Instructions Executed"), hh.index("# Samples")
            cols = [(i, c) for i, c in enumerate(hh) if c.startswith("stall_") and "Not Issued" not in c]
            for r in src[hidx + 1:end]:
                if r and r[0].isdigit():
                    try:
                        n, smp = int(r[ie]), int(r[sm])
                    except ValueError:
                        continue
                    agg.append((n, smp, fname, int(r[0]), r[1].strip()[:100]))
                    tot += n
                    for i, c in cols:
                        try:
                            stalls[c] += int(r[i])
                        except ValueError:
                            pass
        ts = sum(a[1] for a in agg) or 1
        st = sum(stalls.values()) or 1
        f.write("## warp-stall mix (sampled, all warps)\n\n| reason | share |\n|---|---|\n")
        for k, v in stalls.most_common(10):
            f.write(f"| {k} | {100 * v / st:.1f}% |\n")
        f.write(f"\n## hottest source lines (of {tot} executed warp instructions)\n\n| inst % | samples % | line | source |\n|---|---|---|---|\n")
        for n, smp, fname, line, text in sorted(agg, reverse=True)[:top]:
            f.write(f"| {100 * n / max(tot, 1):.1f} | {100 * smp / ts:.1f} | {fname}:{line} | `{text.replace('|', '/')}` |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
