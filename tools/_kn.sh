tools/quick_bench.sh kn_base - --no-newton --no-spot 2>&1 | cut -c1-60
for n in fu3 su1 bu1 sp1 evl azero; do
tools/quick_bench.sh kn_$n tools/variants/lib_$n.so --no-newton --no-spot 2>&1 | cut -c1-60
done
