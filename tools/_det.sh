for rows in 0 4 8 16 32; do
  if [ $rows -gt 0 ]; then export MAF_BAND_ROWS=$rows; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-newton --no-spot > gpurun_out/det_$rows.json 2> gpurun_out/det_$rows.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/det_$rows.json")); print("rows $rows", "atomic ms %.2f"%d["ms_per_step"], d["value_deterministic"])
except Exception as e:
    print("rows $rows FAILED", e); print(open("gpurun_out/det_$rows.err").read()[-800:])
PY
done
