python tools/tune_plan.py --motion EUL --iters 200 --seed 4 --verify-n 1001 > gpurun_out/tune3_eul.log 2>&1
tail -2 gpurun_out/tune3_eul.log
python tools/tune_plan.py --motion ALEV --iters 200 --seed 4 --verify-n 1001 > gpurun_out/tune3_alev.log 2>&1
tail -2 gpurun_out/tune3_alev.log
python tools/tune_plan.py --motion LAG --iters 80 --seed 4 --verify-n 1001 > gpurun_out/tune3_lag.log 2>&1
tail -2 gpurun_out/tune3_lag.log
python tools/tune_plan.py --motion ALEVB --iters 150 --seed 5 --verify-n 1001 > gpurun_out/tune4_alevb.log 2>&1
tail -2 gpurun_out/tune4_alevb.log
