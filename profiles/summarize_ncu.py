#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into JSON: per launch duration, DRAM bytes, pipe/issue
utilisation, occupancy limits and the warp-stall distribution.  Usage:
    python profiles/summarize_ncu.py gpurun_out/X.ncu-rep [out.json]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "duration_us": "gpu__time_duration.sum",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "regs": "launch__registers_per_thread",
    "smem_dyn_B": "launch__shared_mem_per_block_dynamic",
    "dram_read_B": "dram__bytes_read.sum",
    "dram_write_B": "dram__bytes_write.sum",
    "dram_pct_peak": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "pipe_fma_cycles_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "inst_executed": "smsp__inst_executed.sum",
    "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "occ_limit_regs": "launch__occupancy_limit_registers",
    "occ_limit_smem": "launch__occupancy_limit_shared_mem",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k, m in KEYS.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[k] = v * UNIT_SCALE.get(units[i], 1.0) if (k.endswith("_B") or k == "duration_us") else v
        if "dram_read_B" in d and "duration_us" in d:
            d["dram_GBps"] = (d["dram_read_B"] + d.get("dram_write_B", 0.0)) / d["duration_us"] / 1e3
        st = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
                try:
                    st[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(st.values()) or 1.0
        d["stall_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]}
        out.append(d)
    js = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(js + "\n")
    for d in out:
        print("%-44s %8.1f us grid %6d regs %3d  dram %6.0f GB/s (R %.3f W %.3f GB) issue %4.1f%% fma %4.1f%% L2hit %4.1f%%" % (
            d["kernel"][:44], d.get("duration_us", 0), d.get("grid", 0), d.get("regs", 0), d.get("dram_GBps", 0),
            d.get("dram_read_B", 0) / 1e9, d.get("dram_write_B", 0) / 1e9, d.get("issue_active_pct", 0),
            d.get("pipe_fma_cycles_pct", 0), d.get("l2_hit_pct", 0)))
        print("     stalls:", d["stall_pct"])


if __name__ == "__main__":
    main()
