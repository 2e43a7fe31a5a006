python -m pytest tests -m gpu -x -q 2>&1 | tail -3
tools/quick_bench.sh v35 -
tools/quick_bench.sh v35_fu3 tools/variants/lib_fu3.so
tools/quick_bench.sh v35_su1 tools/variants/lib_su1.so
tools/quick_bench.sh v35_lag - --motion LAG
tools/quick_bench.sh v35_lag_bu1 tools/variants/lib_bu1.so --motion LAG
tools/quick_bench.sh v35_lag_su1 tools/variants/lib_su1.so --motion LAG
tools/quick_bench.sh v35_eul - --motion EUL
tools/quick_bench.sh v35_eul_bu1 tools/variants/lib_bu1.so --motion EUL
tools/quick_bench.sh v35_eul_su1 tools/variants/lib_su1.so --motion EUL
tools/ncu_metrics.sh v35
