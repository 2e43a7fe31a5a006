for n in scatter resid geoa geob lin geoa_lin geoa_geob; do
  tools/quick_bench.sh st_$n tools/variants/lib_st_$n.so --no-newton --no-spot 2>&1 | cut -c1-90
done
