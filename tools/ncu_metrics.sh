#!/bin/bash
# usage: tools/ncu_metrics.sh <tag> [lib]   -> gpurun_out/<tag>_metrics.csv (a few unit-utilisation metrics of the area kernel, 301x301 ALEVB)
tag=$1; lib=$2
if [ -n "$lib" ] && [ "$lib" != "-" ]; then export MAF_LIB=$lib; fi
M=gpu__time_duration.sum,sm__cycles_elapsed.avg,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_global_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,lts__t_sectors_srcunit_tex_op_red.sum,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:area_kernel -s 2 -c 1 --csv --log-file gpurun_out/${tag}_metrics.csv python tools/profile_once.py --n 301 > gpurun_out/${tag}_metrics.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${tag}_metrics.csv")) if len(r)>5]
h=rows[0]; i_n=h.index("Metric Name"); i_v=h.index("Metric Value")
print("$tag", {r[i_n].replace("l1tex__data_pipe_lsu_wavefronts","wf").replace(".sum","").replace("smsp__",""): r[i_v] for r in rows[1:]})
PY
