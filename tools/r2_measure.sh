#!/bin/bash
# round-2 measurement set on ONE B200 (run under gpurun from the repo root); outputs in gpurun_out/
set -x
python bench.py > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err
for m in LAG EUL ALEV; do
  python bench.py --motion $m --steps 5 --no-cpu --no-newton --no-spot > gpurun_out/r2_final_n1_$m.json 2> gpurun_out/r2_final_n1_$m.err
done
python bench.py --patch-n 2049 --steps 3 --no-cpu --no-newton --no-spot > gpurun_out/r2_final_n1_2049.json 2> gpurun_out/r2_final_n1_2049.err
# launch list of the bench command (share of each kernel in a step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launch_list_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-newton --no-spot > gpurun_out/r2_launch_list.log 2>&1
# full capture of the dominant kernel at the bench workload
ncu --set full --import-source on --clock-control none -k regex:area_kernel -s 2 -c 1 -o gpurun_out/r2_area_1001 -f \
    python tools/profile_once.py --n 1001 > gpurun_out/r2_ncu_full.log 2>&1
# sanitizer
for tool in memcheck racecheck synccheck; do
  echo "## $tool" >> gpurun_out/r2_sanitizer.txt
  compute-sanitizer --tool $tool python tools/sanitize_once.py 2>&1 | grep -E "^ok|SUMMARY|ERROR|hazard" | head -20 >> gpurun_out/r2_sanitizer.txt
done
ls -la gpurun_out/r2_area_1001.ncu-rep
