#!/usr/bin/env python3
"""One tracked JSON record per captured kernel launch of an .ncu-rep -- the numbers bench.py reads back for
`roofline.traffic` and the hardware view (instruction count, pipe utilisation).

    python tools/ncu_json.py rep.ncu-rep [kernel-name-substring] > profiles/x.json
"""
import csv
import io
import json
import subprocess
import sys

FIELDS = {
    "duration_us": "gpu__time_duration.sum",
    "warp_inst": "smsp__inst_executed.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "fma_inst_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_bytes": "lts__t_bytes.sum",
    "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size", "block": "launch__block_size",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "stall_math_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_mio_throttle": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "stall_lg_throttle": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "stall_branch": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else None
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    out = []
    for r in rows[2:]:
        if want and want not in r[ki]:
            continue
        rec = {"kernel": r[ki][:160], "source": f"ncu --set full --clock-control none ({rep.split('/')[-1]})"}
        for k, m in FIELDS.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if k.endswith("_bytes") or k == "duration_us":
                    v *= SCALE.get(u, 1.0)
                rec[k] = v
        if "dram_read_bytes" in rec:
            rec["dram_bytes"] = rec["dram_read_bytes"] + rec.get("dram_write_bytes", 0.0)
        out.append(rec)
    print(json.dumps(out[0] if len(out) == 1 else out, indent=1))


if __name__ == "__main__":
    main()
