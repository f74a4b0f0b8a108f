#!/usr/bin/env python
"""Condense ncu outputs into small tracked files under profiles/.

  ncu_summary.py launches <launches.csv> <out.json>     per-kernel totals / share of the profiled run
  ncu_summary.py full <raw.csv> <out.json> [kernel-substring]   key counters of one `--set full` capture
                 (raw.csv = `ncu -i x.ncu-rep --page raw --csv`)
  ncu_summary.py traffic <raw.csv> <fusion> <users_per_launch> <items_per_rank> <source-label>
                 add / replace the DRAM bytes of that launch shape in profiles/ncu_traffic.json (bench.py: roofline.traffic)
"""
import csv
import json
import re
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "sm__clocks_per_second", "gpc__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg.per_second",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void\s+", "", name)
    return name.strip()[-90:]


def launches(path, out):
    tot = defaultdict(lambda: [0, 0.0])
    with open(path, newline="") as f:
        rd = csv.reader(l for l in f if l.startswith('"'))
        hdr = next(rd)
        ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
        for r in rd:
            if len(r) > vi and r[mi] == "gpu__time_duration.sum":
                t = tot[short(r[ki])]
                t[0] += 1
                t[1] += float(r[vi].replace(",", ""))
    total = sum(v[1] for v in tot.values())
    rows = sorted(({"kernel": k, "launches": n, "total_ms": ns / 1e6, "avg_ms": ns / 1e6 / n, "share": ns / total}
                   for k, (n, ns) in tot.items()), key=lambda r: -r["total_ms"])
    json.dump({"source": path, "note": "ncu --metrics gpu__time_duration.sum --clock-control none: serialised, cold-cache "
               "per-launch times of the whole process (workload generation included); shares, not absolutes",
               "total_ms": total / 1e6, "kernels": rows[:25]}, open(out, "w"), indent=1)
    for r in rows[:12]:
        print(f"{r['share']:.4f} {r['total_ms']:10.3f} ms x{r['launches']:<4d} {r['kernel']}")


def full(path, out, sub=None):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    res = []
    for r in rows[2:]:
        if sub and sub not in r[ki]:
            continue
        d = {"kernel": short(r[ki])}
        for h, u, v in zip(hdr, units, r):
            base = h.split("TriageCompute.")[-1]
            if base in KEYS or h in KEYS:
                try:
                    d[base] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    pass
        res.append(d)
    json.dump({"source": path, "captures": res}, open(out, "w"), indent=1)
    for d in res:
        for k, v in d.items():
            print(k, v)


def traffic(path, fusion, users, items, label):
    from pathlib import Path
    rows = list(csv.reader(open(path, newline="")))
    hdr, r = rows[0], rows[2]
    d = {h: v for h, v in zip(hdr, r)}
    u = {h: v for h, v in zip(hdr, rows[1])}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = sum(float(d[k].replace(",", "")) * mult[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    out = Path(__file__).resolve().parent.parent / "profiles" / "ncu_traffic.json"
    j = json.loads(out.read_text()) if out.exists() else {"note": "", "captures": []}
    j["captures"] = [c for c in j["captures"] if not (c["fusion"] == fusion and c["users_per_launch"] == int(users) and c["items_per_rank"] == int(items))]
    j["captures"].append({"fusion": fusion, "users_per_launch": int(users), "items_per_rank": int(items),
                          "dram_bytes_per_launch": int(tot), "source": label})
    out.write_text(json.dumps(j, indent=1))
    print(fusion, users, items, int(tot))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "traffic":
        traffic(*sys.argv[2:7])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
