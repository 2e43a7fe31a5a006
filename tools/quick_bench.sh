#!/bin/bash
# usage: tools/quick_bench.sh <tag> [lib] [extra bench args]  -> prints one summary line, writes gpurun_out/<tag>.json
tag=$1; lib=$2; shift; shift
if [ -n "$lib" ] && [ "$lib" != "-" ]; then export MAF_LIB=$lib; fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$tag.json"))
    print("$tag", "Melem/s %.2f"%d["value"], "kernel_ms %.2f"%d["roofline"]["kernel_ms"], "fp64 frac %.3f"%d["roofline"]["fp64"]["frac"], d["config"]["kernel"]["ctas_per_sm"], "ctas/sm", d["roofline"]["other_kernels_ms"])
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/$tag.err").read()[-1500:])
PY
