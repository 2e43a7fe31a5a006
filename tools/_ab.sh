python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for m in ALEVB LAG EUL ALEV; do
tools/quick_bench.sh pm_$m - --no-newton --no-spot --motion $m 2>&1 | cut -c1-80
done
