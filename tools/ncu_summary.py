"""Text summary of one kernel of an .ncu-rep (the form kept under profiles/): a fixed list of metrics, the stall
reasons per issued instruction and the DRAM traffic as JSON.

usage: python tools/ncu_summary.py <report.ncu-rep> <out.txt> [traffic.json] [header line ...]
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads",
    "sass__inst_executed_shared_stores",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed_op_global_red.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    tjson = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3].endswith(".json") else None
    header = sys.argv[4:] if tjson else sys.argv[3:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    m = {n: (u, v) for n, u, v in zip(names, units, vals)}
    lines = [f"# {h}" for h in header]
    lines.append(f"# kernel: {m.get('Kernel Name', ('', '?'))[1]}")
    for k in METRICS:
        if k in m:
            lines.append(f"{k:85s} {m[k][0]:15s} {m[k][1]}")
    for n in sorted(m):   # whatever form the FP64 thread-instruction counters were collected in
        if n.startswith("smsp__sass_thread_inst_executed_op_d") and ".sum" in n and n not in METRICS:
            lines.append(f"{n:85s} {m[n][0]:15s} {m[n][1]}")
    issued = None
    for k in ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",):
        if k in m:
            issued = float(m[k][1].replace(",", ""))
    lines.append("--- stall reasons per issue (smsp__average_warps_issue_stalled_*_per_issue_active.ratio)")
    st = []
    for n, (u, v) in m.items():
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and "not_issued" not in n:
            try:
                st.append((float(v.replace(",", "")), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    for v, n in sorted(st, reverse=True):
        lines.append(f"  {n:30s} {v:.2f}")
    open(out, "w").write("\n".join(lines) + "\n")

    def num(k):
        u, v = m[k]
        x = float(v.replace(",", ""))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
                 "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
        return x * scale.get(u, 1.0)
    if tjson:
        json.dump({"dram_read": num("dram__bytes_read.sum"), "dram_write": num("dram__bytes_write.sum"),
                   "time": num("gpu__time_duration.sum")}, open(tjson, "w"))
    print(open(out).read())


if __name__ == "__main__":
    main()
