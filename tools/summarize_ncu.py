#!/usr/bin/env python
"""Turn ncu exports into the small tracked summaries under profiles/.

  ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/summarize_ncu.py full raw.csv out.json
  python tools/summarize_ncu.py launches launches.csv out.json      (gpu__time_duration pass)
"""
import collections
import csv
import json
import re
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]


def full(path, out, what=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        if what:
            d["what"] = what
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        st = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
                try:
                    st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[i])
                except ValueError:
                    pass
        tot = sum(st.values()) or 1.0
        d["stall_samples_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]}
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        name = row["Kernel Name"]
        m = re.search(r"ConvCfg<([^>]*)>", name)
        key = "conv1d_f32<" + re.sub(r"\(int\)", "", m.group(1)) + ">" if m else name.split("(")[0][:70]
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    res = {"total_ms": tot, "note": "per-launch device time under ncu: cold caches, serialised -- compare shares",
           "kernels": [dict(kernel=k, launches=n, ms=ms, share=ms / tot) for k, (n, ms) in
                       sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    json.dump(res, open(out, "w"), indent=1)
    for k in res["kernels"][:15]:
        print(f"{k['ms']:9.3f} ms {100 * k['share']:5.1f}% n={k['launches']:4d} {k['kernel']}")


def dram(path, out, batch="32"):
    """single-pass ncu of one call's tensor-core conv launches (tools/gpu_ncu.sh): DRAM bytes per launch of the generator's
    ResBlock convs, in launch order (see tools/ncu_target.py for the order)."""
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1.0)
        per.setdefault(i, {"kernel": row["Kernel Name"].split("(")[0]})[row["Metric Name"]] = v
    ids = sorted(per)
    assert len(ids) == 167, f"expected the 167 tensor-core conv launches of one call, got {len(ids)}"
    stages = {"stage0_C256": range(98, 116), "stage1_C128": range(117, 135), "stage2_C64": range(136, 151), "stage3_C32": range(152, 167)}
    res = {"what": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none (single pass, no "
                   f"replay), one convert call at batch {batch} x 10 s, f16x3; the 66 launches of the 72 generator ResBlock convs "
                   f"(the six k = 3 pairs of stages 2-3 are fused)", "stages": {}}
    tot_b, tot_ms = 0.0, 0.0
    for name, rng in stages.items():
        b = sum(per[ids[k]]["dram__bytes_read.sum"] + per[ids[k]]["dram__bytes_write.sum"] for k in rng)
        ms = sum(per[ids[k]]["gpu__time_duration.sum"] for k in rng)
        kern = sorted({per[ids[k]]["kernel"] for k in rng})
        res["stages"][name] = {"launches": len(rng), "dram_bytes": b, "ms_under_ncu": ms, "kernels": kern}
        tot_b += b
        tot_ms += ms
    n_launch = sum(len(r) for r in stages.values())
    res["resblock_launches"] = n_launch
    res["dram_bytes_per_launch_avg"] = tot_b / n_launch
    res["dram_bytes_total"] = tot_b
    res["ms_under_ncu_total"] = tot_ms
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"full": full, "launches": launches, "dram": dram}[sys.argv[1]](*sys.argv[2:])
