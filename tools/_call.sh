python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err; tail -c 600 gpurun_out/r1f_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1f_ref.json 2> gpurun_out/r1f_ref.err; tail -c 400 gpurun_out/r1f_ref.json
tools/quick_bench.sh v36 -
tools/quick_bench.sh v36_lag - --motion LAG
tools/quick_bench.sh v36_eul - --motion EUL
tools/quick_bench.sh v36_alev - --motion ALEV
tools/quick_bench.sh v36_det - --scatter deterministic
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r1f_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:area_kernel -s 2 -c 1 -f -o gpurun_out/v36_1001 python tools/profile_once.py --n 1001 > gpurun_out/v36_1001_ncu.log 2>&1
tail -2 gpurun_out/v36_1001_ncu.log
tools/ncu_metrics.sh v36
