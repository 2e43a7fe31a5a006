python -m pytest tests/test_gpu_parity.py tests/test_gpu_strips.py -m gpu -x -q -k "determin or banded or strip or matches_oracle" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu --no-newton --no-spot > gpurun_out/det_pc2.json 2> gpurun_out/det_pc2.err
python -c "
import json; d=json.load(open('gpurun_out/det_pc2.json')); print(d['value'], d['value_deterministic'], d['config']['setup_s'])"
python bench.py --patch-n 2049 --steps 3 --warmup 2 --no-cpu --no-newton --no-spot > gpurun_out/det_pc2_2049.json 2> gpurun_out/det_pc2_2049.err
python -c "
import json; d=json.load(open('gpurun_out/det_pc2_2049.json')); print(d['value'], d['value_deterministic'], d['e2e']['value'])"
