set -x
for m in LAG EUL ALEV; do
  python bench.py --motion $m --steps 5 --no-cpu --no-newton --no-spot > gpurun_out/r2_end_n1_$m.json 2> gpurun_out/r2_end_n1_$m.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launch_list_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-newton --no-spot > gpurun_out/r2_launch_list.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:area_kernel -s 2 -c 1 -o gpurun_out/r2_area_1001 -f \
    python tools/profile_once.py --n 1001 > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/r2_area_1001.ncu-rep
python bench.py > gpurun_out/r2_end_n1.json 2> gpurun_out/r2_end_n1.err
