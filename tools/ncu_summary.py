#!/usr/bin/env python
"""Turns an `ncu --set full --import-source on` report into the markdown summary committed under profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.md
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep --counters C2_strict profiles/rNN_<what>.md K T pipe|mono
      (additionally records warp-instructions / DRAM bytes per launch of the first captured launch under that key in
       profiles/kernel_counters.json -- with the K, T and kernel variant they were captured at -- which bench.py reads
       for roofline.issue / roofline.traffic and uses as they are ONLY for that configuration)

Runs here (no GPU needed: `ncu -i` only reads the report).
"""
from __future__ import annotations

import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum.per_cycle_elapsed", "sm__inst_executed.sum.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sector_op_read_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sectors.sum.per_second",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def update_counters(raw, key, source, K=None, T=None, variant=None):
    import json
    import os
    hdr = raw[0]
    d = dict(zip(hdr, raw[2]))
    f = lambda k: float(d[k].replace(",", "")) if d.get(k) else 0.0                       # noqa: E731
    units = dict(zip(hdr, raw[1]))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = sum(f(k) * scale.get(units.get(k, "byte"), 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "kernel_counters.json")
    allc = json.load(open(path)) if os.path.exists(path) else {}
    allc[key] = {"K": K, "T": T, "variant": variant,
                 "warp_inst_per_launch": f("smsp__inst_executed.sum"), "dram_bytes_per_launch": dram,
                 "l2_sectors_per_launch": f("lts__t_sectors.sum"), "l2_hit_pct": f("lts__t_sector_hit_rate.pct"),
                 "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "l1_hit_pct": f("l1tex__t_sector_hit_rate.pct"), "kernel": d.get("Kernel Name"),
                 "source": f"{source} (ncu --set full, first captured launch)"}
    json.dump(allc, open(path, "w"), indent=1)


def main():
    rep = sys.argv[1]
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    if len(sys.argv) >= 5 and sys.argv[2] == "--counters":
        extra = sys.argv[5:8]
        update_counters(raw, sys.argv[3], sys.argv[4], int(extra[0]) if len(extra) > 0 else None,
                        int(extra[1]) if len(extra) > 1 else None, extra[2] if len(extra) > 2 else None)
    print(f"# ncu summary of `{rep.split('/')[-1]}`\n")
    print("`ncu --set full --clock-control none` (tools/prof_all.sh; one replayed launch per row; times under the "
          "profiler are cold-cache and serialised and are NOT bench numbers; the SASS section is present when the "
          "capture was taken with `--import-source on`).\n")
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        print(f"## launch {d.get('ID')}: `{d.get('Kernel Name')}` grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
        print("| metric | value | unit |\n|---|---|---|")
        u = dict(zip(hdr, units))
        for k in RAW_KEYS:
            if k in d and d[k] != "":
                print(f"| {k} | {d[k]} | {u[k]} |")
        st = [(h, float(d[h])) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h]]
        st.sort(key=lambda x: -x[1])
        print("\nWarp states per issue slot (warps per issue-active cycle; `selected` = issuing):\n")
        print("| state | warps |\n|---|---|")
        for h, v in st[:9]:
            print(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.3f} |")
        print()
    src = ncu_csv(rep, "source")
    starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    if not starts:
        return
    s0 = starts[0]
    end = starts[1] if len(starts) > 1 else len(src)
    h = src[s0 + 1]
    ix = {n: i for i, n in enumerate(h)}
    data = src[s0 + 2:end]
    reasons = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    agg = collections.Counter()
    ops = collections.Counter()
    for r in data:
        for n in reasons:
            agg[n] += int(r[ix[n]] or 0)
        toks = r[ix["Source"]].split()
        if toks:
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
    print(f"## SASS-level sampling of the first launch ({len(data)} instructions, {tot} samples)\n")
    print("| stall reason | samples | share |\n|---|---|---|")
    for n, v in agg.most_common(10):
        print(f"| {n} | {v} | {100.0 * v / max(1, sum(agg.values())):.1f} % |")
    nexec = sum(ops.values())
    print(f"\nExecuted warp-instructions by opcode ({nexec} total):\n")
    print("| opcode | warp-instructions | share |\n|---|---|---|")
    for op, v in ops.most_common(18):
        print(f"| {op} | {v} | {100.0 * v / max(1, nexec):.1f} % |")
    print("\nTop 25 instructions by samples:\n")
    print("| # | SASS | samples | executed | dominant stalls |\n|---|---|---|---|---|")
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:25]
    for i in sorted(top):
        r = data[i]
        rs = sorted(((n, int(r[ix[n]] or 0)) for n in reasons), key=lambda x: -x[1])[:2]
        rs = ", ".join(f"{n.replace('stall_', '')} {v}" for n, v in rs if v)
        print(f"| {i} | `{r[ix['Source']].strip()[:64]}` | {r[ix['# Samples']]} | {r[ix['Instructions Executed']]} | {rs} |")


if __name__ == "__main__":
    main()
