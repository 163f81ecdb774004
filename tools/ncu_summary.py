#!/usr/bin/env python
"""tools/ncu_summary.py -- condense an `ncu --set full` report into the table kept under profiles/.

  ncu -i gpurun_out/prof_v11.ncu-rep --page raw --csv > /tmp/raw.csv
  python tools/ncu_summary.py /tmp/raw.csv "header comment" > profiles/r1_v11_ncu_full_summary.csv
  python tools/ncu_summary.py /tmp/raw.csv --traffic       # JSON of DRAM bytes per launch per kernel

One column per distinct kernel (first launch of each kept), one row per metric of METRICS.
"""
import csv
import json
import sys

METRICS = """launch__grid_size launch__block_size launch__registers_per_thread gpu__time_duration.sum
dram__bytes_read.sum dram__bytes_write.sum sm__throughput.avg.pct_of_peak_sustained_elapsed
smsp__issue_active.avg.pct_of_peak_sustained_active sm__warps_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
launch__occupancy_limit_shared_mem launch__occupancy_limit_registers""".split()


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    kcol = col["Kernel Name"]
    kernels, seen = [], set()
    for r in rows[hdr + 2:]:
        if len(r) <= kcol or r[kcol] in seen:
            continue
        seen.add(r[kcol])
        kernels.append(r)
    if len(sys.argv) > 2 and sys.argv[2] == "--traffic":
        out = {}
        for r in kernels:
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            out[r[kcol]] = int(rd + wr)
        print(json.dumps(out, indent=1))
        return
    w = csv.writer(sys.stdout)
    if len(sys.argv) > 2:
        w.writerow(["# " + sys.argv[2]])
    w.writerow(["metric", "unit"] + [r[kcol].split("(")[0] for r in kernels])
    for m in METRICS:
        if m not in col:
            continue
        w.writerow([m, units[col[m]]] + [r[col[m]] for r in kernels])


if __name__ == "__main__":
    main()
