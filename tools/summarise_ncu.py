"""Turns an .ncu-rep (ncu --set full --import-source on) into the JSON summaries kept under profiles/.
usage: python tools/summarise_ncu.py REPORT.ncu-rep OUT.json [kernel-substring] [algorithmic_bytes_per_launch]"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if want and want not in name:
            continue
        j = {"kernel": name, "report": rep.split("/")[-1]}
        for k in KEYS:
            if k in hdr:
                j[k] = {"value": r[hdr.index(k)], "unit": units[hdr.index(k)]}
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
                j[h] = {"value": r[i], "unit": units[i]}
        rd, wr = j.get("dram__bytes_read.sum"), j.get("dram__bytes_write.sum")
        if rd and wr:
            j["traffic_bytes_per_launch"] = to_bytes(rd["value"], rd["unit"]) + to_bytes(wr["value"], wr["unit"])
        if len(sys.argv) > 4:
            j["algorithmic_bytes_per_launch"] = float(sys.argv[4])
        res.append(j)
    json.dump(res[0] if len(res) == 1 else res, open(out, "w"), indent=1)
    print("wrote", out, len(res), "kernel(s)")


if __name__ == "__main__":
    main()
