#!/bin/bash
# Timing-only builds that drop one piece of the area kernel each (WRONG results; -DMAF_STUB_* in maf_element.cuh /
# maf_api.cu) and their kernel times on the bench workload: what each piece costs on the critical path, and which
# pieces are co-critical (profiles/r2_variants.md). Run the build part here, the timing part under gpurun:
#   tools/stub_sweep.sh build            -> tools/variants/lib_st_<name>.so
#   gpurun -- 'tools/stub_sweep.sh time [MOTION]'
set -e
VARIANTS=("scatter -DMAF_STUB_SCATTER" "red -DMAF_STUB_RED" "resid -DMAF_STUB_RESIDUAL" "geoa -DMAF_STUB_GEO_A"
          "geob -DMAF_STUB_GEO_B" "lin -DMAF_STUB_LIN" "geoa_lin -DMAF_STUB_GEO_A -DMAF_STUB_LIN"
          "geoa_geob -DMAF_STUB_GEO_A -DMAF_STUB_GEO_B" "gauss -DMAF_STUB_GAUSS" "tangent -DMAF_STUB_TANGENT"
          "azero -DMAF_STUB_AZERO")
if [ "$1" = "build" ]; then
  for v in "${VARIANTS[@]}"; do set -- $v; n=$1; shift; bash tools/build_variant.sh st_$n "$@"; done
else
  motion=${2:-ALEVB}
  tools/quick_bench.sh st_base - --no-newton --no-spot --motion $motion | cut -c1-70
  for v in "${VARIANTS[@]}"; do set -- $v
    tools/quick_bench.sh st_$1 tools/variants/lib_st_$1.so --no-newton --no-spot --motion $motion | cut -c1-70
  done
fi
